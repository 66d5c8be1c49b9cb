// CPU baseline of SURVEY.md 8(d): the reference's OWN JavaScript (jsfft + src/extractors, unmodified, loaded
// from a checkout of kirbysayshi/meyda) timed on worker_threads, one clip range per worker.
//
//   node baseline/node/ref_worker.js --ref /path/to/meyda [--threads T] [--clips C] [--seconds S]
//        [--bufferSize 2048] [--hop 512] [--features all|name,name,...] [--window hanning]
//
// prints ONE JSON line {"impl": "reference", "metric": "feature frames/sec ...", "value": ..., "cpu_baseline":
// {"kind": "reference", "cores": T, ...}} in bench.py's format.
//
// STATUS: Node.js is absent from the build image and from the GPU boxes, so the worker_threads / require() shell
// of this file has never been executed.  The numeric core between the "ES5 CORE" markers HAS been executed: it
// is plain ES5, and tests/test_js_pin.py runs exactly that text under oracle/minijs.py against the reference
// sources and the committed golden vectors.  bench.py --impl reference reports the C restatement instead
// (cpu_baseline.kind "port").
//
// The core performs the intended per-buffer sequence of src/meyda.js:69-91 with the wiring fixes of SURVEY.md
// 2.3 (a fresh zero-imaginary ComplexArray transformed per frame, `buffer` = the raw frame, perceptual* reaching
// loudness through m.featureExtractors.loudness, mfcc's free `audioContext`, the free global `µ`).  Every number
// is computed by the reference's code; window and bark tables are built once, as `new Meyda(...)` does.

// ---- BEGIN ES5 CORE
function makeReferencePath(env, N, sampleRate, windowName, names) {
  // env: {ComplexArray, extractors: {name: fn}, computeAmplitude, computeHanning, computeHamming, computeWindow,
  //       computeBarkScale}  (the compute* bodies of src/meyda.js:104-182 lifted out of the class, verbatim)
  var m = {signal: null, audioContext: {sampleRate: sampleRate}, featureExtractors: {}};
  m.barkScale = env.computeBarkScale.call(m, N, sampleRate);
  m.hanning = env.computeHanning.call(m, N);
  m.hamming = env.computeHamming.call(m, N);
  m.ampSpectrum = new Float32Array(N / 2);
  var loud = env.extractors.loudness({NUM_BARK_BANDS: 24, barkScale: m.barkScale,
                                      normalisedSpectrum: m.ampSpectrum, sampleRate: sampleRate});
  m.featureExtractors.loudness = function(bufferSize, mm) { return loud.process(); };
  var ComplexArray = env.ComplexArray;
  return {
    m: m,
    loudness: loud,
    frame: function(signal) {
      m.signal = signal;
      var windowedSignal = env.computeWindow.call(m, signal, windowName);
      var data = new ComplexArray(N);
      data.map(function(value, i, n) { value.real = windowedSignal[i]; });
      var spec = data.FFT();
      m.complexSpectrum = spec;
      env.computeAmplitude.call(m, spec, m.ampSpectrum, N);
      var results = {};
      for (var x = 0; x < names.length; x++) {
        var name = names[x];
        if (name == "buffer") results[name] = m.signal;
        else if (name == "loudness") results[name] = loud.process();
        else results[name] = env.extractors[name](N, m);
      }
      return results;
    }
  };
}

function framesOf(len, N, hop) { return len < N ? 0 : Math.floor((len - N) / hop) + 1; }

function runClips(path, clips, N, hop) {
  // clips: array of Float32Array.  Returns the number of frames processed; results are dropped like a callback
  // that does nothing (the reference returns aliases of reused buffers anyway).
  var frames = 0, sink = 0;
  for (var c = 0; c < clips.length; c++) {
    var clip = clips[c], nf = framesOf(clip.length, N, hop);
    for (var f = 0; f < nf; f++) {
      var r = path.frame(clip.subarray(f * hop, f * hop + N));
      if (r.rms !== undefined) sink += r.rms;
      frames++;
    }
  }
  return {frames: frames, sink: sink};
}
// ---- END ES5 CORE

var ALL = ["buffer", "rms", "energy", "zcr", "complexSpectrum", "amplitudeSpectrum", "powerSpectrum",
           "spectralCentroid", "spectralFlatness", "spectralSlope", "spectralRolloff", "spectralSpread",
           "spectralSkewness", "spectralKurtosis", "loudness", "perceptualSpread", "perceptualSharpness", "mfcc"];

function loadReference(ref) {
  var fs = require('fs'), path = require('path');
  var ComplexArray = require(path.join(ref, 'lib/jsfft/complex_array')).ComplexArray;
  require(path.join(ref, 'lib/jsfft/fft'));  // decorates ComplexArray.prototype with FFT
  var utils = require(path.join(ref, 'src/utils'));
  global['µ'] = utils['µ'];  // free global in spectralCentroid / Spread / Skewness / Kurtosis.js
  var env = {ComplexArray: ComplexArray, extractors: {}};
  ALL.concat(["loudness"]).forEach(function(n) {
    if (n != "buffer") env.extractors[n] = require(path.join(ref, 'src/extractors', n));  // buffer.js is empty
  });
  // src/meyda.js is an ES6 class wired to window / Web Audio: lift the five compute* method bodies, verbatim
  var src = fs.readFileSync(path.join(ref, 'src/meyda.js'), 'utf8');
  ["computeAmplitude", "computeHamming", "computeHanning", "computeWindow", "computeBarkScale"].forEach(function(name) {
    var mt = new RegExp("\\n\\t+" + name + "\\(([^)]*)\\)\\s*\\{").exec(src);
    var i = mt.index + mt[0].length, start = i, depth = 1;
    while (depth) { var ch = src[i++]; if (ch == "{") depth++; else if (ch == "}") depth--; }
    env[name] = new Function(mt[1], src.slice(start, i - 1));
  });
  return env;
}

function synthClip(index, len, sampleRate) {
  // white noise + three sines, amplitude inside (-0.85, 0.85); xorshift32 keyed by the clip index
  var s = (0x4D455944 ^ Math.imul(index + 1, 0x9E3779B1)) >>> 0 || 1;
  function u() { s ^= s << 13; s >>>= 0; s ^= s >>> 17; s ^= s << 5; s >>>= 0; return s / 4294967296; }
  var f = [], ph = [];
  for (var p = 0; p < 3; p++) { f.push(55 * Math.pow(15000 / 55, u())); ph.push(2 * Math.PI * u()); }
  var x = new Float32Array(len);
  for (var i = 0; i < len; i++) {
    var v = 0.5 * (u() - 0.5);
    for (p = 0; p < 3; p++) v += 0.2 * Math.sin(2 * Math.PI * f[p] * i / sampleRate + ph[p]);
    x[i] = v;
  }
  return x;
}

function parseArgs(argv) {
  var a = {ref: process.env.MEYDA_REF || '/root/reference', threads: require('os').cpus().length, clips: 64,
           seconds: 30, bufferSize: 2048, hop: 512, features: 'all', window: 'hanning', sampleRate: 44100};
  for (var i = 2; i + 1 < argv.length; i += 2) {
    var k = argv[i].replace(/^--/, '');
    a[k] = (typeof a[k] == 'number') ? Number(argv[i + 1]) : argv[i + 1];
  }
  a.names = a.features == 'all' ? ALL : a.features.split(',');
  return a;
}

if (typeof require != 'undefined' && typeof module != 'undefined' && require.main === module) {
  var wt = require('worker_threads');
  if (wt.isMainThread) {
    var args = parseArgs(process.argv), done = 0, frames = 0, t0 = 0, ready = 0, workers = [];
    for (var t = 0; t < args.threads; t++) {
      var lo = Math.floor(args.clips * t / args.threads), hi = Math.floor(args.clips * (t + 1) / args.threads);
      var w = new wt.Worker(__filename, {workerData: {args: args, lo: lo, hi: hi}});
      workers.push(w);
      w.on('message', function(msg) {
        if (msg.ready) {  // every worker has its clips and tables: start the clock, release them together
          if (++ready == args.threads) { t0 = process.hrtime.bigint(); workers.forEach(function(x) { x.postMessage('go'); }); }
          return;
        }
        frames += msg.frames;
        if (++done == args.threads) {
          var sec = Number(process.hrtime.bigint() - t0) / 1e9, v = frames / sec;
          var sample = args.clips + " clips x " + args.seconds + " s, bufferSize " + args.bufferSize + " hop " + args.hop +
                       ", features " + args.features + ", " + frames + " frames in " + sec.toFixed(2) + " s";
          console.log(JSON.stringify({impl: "reference", metric: "feature frames/sec (full set, N=2048)", value: v,
            unit: "frames/s", n_gpus: 0, higher_is_better: true, dtype: "f64", data: "synthetic",
            cpu_baseline: {value: v, unit: "frames/s", cores: args.threads, kind: "reference", sample: sample,
                           node: process.version, cpu: require('os').cpus()[0].model},
            e2e: {value: v, unit: "frames/s", h2d_bytes_per_step: 0, d2h_bytes_per_step: 0}}));
          process.exit(0);
        }
      });
    }
  } else {
    var d = wt.workerData, A = d.args;
    global.audioContext = {sampleRate: A.sampleRate};  // free global in mfcc.js:20,37
    var path_ = makeReferencePath(loadReference(A.ref), A.bufferSize, A.sampleRate, A.window, A.names);
    var clips = [];
    for (var c = d.lo; c < d.hi; c++) clips.push(synthClip(c, Math.floor(A.seconds * A.sampleRate), A.sampleRate));
    runClips(path_, clips.slice(0, 1).map(function(x) { return x.subarray(0, 16 * A.bufferSize); }), A.bufferSize, A.hop);  // JIT warm-up
    wt.parentPort.once('message', function() { wt.parentPort.postMessage(runClips(path_, clips, A.bufferSize, A.hop)); });
    wt.parentPort.postMessage({ready: true});
  }
}
