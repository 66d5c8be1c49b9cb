// meyda_b200.js -- JavaScript facade over the N-API addon: the reference's
// extractor API (feature names, featureInfo, get([...]) shapes, per-buffer
// callback) served from one batched GPU call.
//
// NOT RUN IN THIS REPO'S IMAGE (no Node).  meyda_b200/meyda.py is the tested
// mirror of this file; keep the two in step.
'use strict'
const native = require('./build/Release/meyda_b200.node')

// src/feature-info.js:3-64
const featureInfo = {
  buffer: {type: 'array'}, rms: {type: 'number'}, energy: {type: 'number'}, zcr: {type: 'number'},
  complexSpectrum: {type: 'multipleArrays', arrayNames: {1: 'real', 2: 'imag'}},
  amplitudeSpectrum: {type: 'array'}, powerSpectrum: {type: 'array'},
  spectralCentroid: {type: 'number'}, spectralFlatness: {type: 'number'}, spectralSlope: {type: 'number'},
  spectralRolloff: {type: 'number'}, spectralSpread: {type: 'number'}, spectralSkewness: {type: 'number'},
  spectralKurtosis: {type: 'number'},
  loudness: {type: 'multipleArrays', arrayNames: {1: 'total', 2: 'specific'}},
  perceptualSpread: {type: 'number'}, perceptualSharpness: {type: 'number'}, mfcc: {type: 'array'}
}
const FEATURES = Object.keys(featureInfo)
const FIELD = {  // feature -> [mb_outputs field, per-frame length(N, out)]
  buffer: ['buffer', N => N], rms: ['rms'], energy: ['energy'], zcr: ['zcr'],
  amplitudeSpectrum: ['amplitude_spectrum', N => N / 2], powerSpectrum: ['power_spectrum', N => N / 2],
  spectralCentroid: ['spectral_centroid'], spectralFlatness: ['spectral_flatness'], spectralSlope: ['spectral_slope'],
  spectralRolloff: ['spectral_rolloff'], spectralSpread: ['spectral_spread'], spectralSkewness: ['spectral_skewness'],
  spectralKurtosis: ['spectral_kurtosis'], perceptualSpread: ['perceptual_spread'],
  perceptualSharpness: ['perceptual_sharpness'], mfcc: ['mfcc', (N, out) => out.numMfccCoefficients || 13]
}

// src/utils.js:13-19
function isPowerOfTwo (num) {
  while (((num % 2) === 0) && num > 1) num /= 2
  return num === 1
}

function splitFeatures (features) {
  if (typeof features === 'string') return [features]
  if (typeof features === 'object' && features !== null) {
    return features.filter(f => {  // src/meyda.js:248-254: unknown names are reported and skipped
      if (featureInfo[f]) return true
      console.error(new TypeError('unknown feature ' + f))
      return false
    })
  }
  throw new Error('Invalid Feature Format')  // src/meyda.js:259
}

// Plans are the tables `new Meyda(...)` precomputes (src/meyda.js:44-48) resident on a GPU: creating one costs a few
// hundred microseconds and a handful of device allocations, so they are kept per parameter set (the reference builds
// its tables once per Meyda object, never per buffer).  clearPlans() releases them.
const planCache = new Map()
function planFor (opts, N, features) {
  const featureMask = features.reduce((m, f) => m | (1 << FEATURES.indexOf(f)), 0)
  const o = {
    bufferSize: N, hop: opts.hop || N, sampleRate: opts.sampleRate || 44100, featureMask,
    window: {hanning: 0, hamming: 1, blackman: 2}[opts.windowingFunction || 'hanning'] || 0, device: opts.device || 0,
    flags: opts.flags || 0,
    numBarkBands: opts.numBarkBands || 0, numMelFilters: opts.numMelFilters || 0,  // 0: the reference's 24 / 26 / 13 / 0.99
    numMfccCoefficients: opts.numMfccCoefficients || 0, rolloffFraction: opts.rolloffFraction || 0
  }
  const key = JSON.stringify(o)
  let plan = planCache.get(key)
  if (!plan) { plan = native.createPlan(o); planCache.set(key, plan) }
  return plan
}
function clearPlans () { planCache.forEach(p => native.destroyPlan(p)); planCache.clear() }

function wrapResult (out, N, features, callback) {
  const result = {
    features, arrays: out, totalFrames: Number(out.totalFrames),
    value: (i, f) => frameValue(out, N, i, f),
    frame: i => Object.fromEntries(features.map(f => [f, frameValue(out, N, i, f)]))
  }
  if (typeof callback === 'function') {  // the buffer-by-buffer contract, src/meyda.js:87-89 (on the JS thread)
    for (let i = 0; i < result.totalFrames; i++) callback(result.frame(i))
  }
  return result
}

function packClips (clips) {
  const list = Array.isArray(clips) ? clips : [clips]
  const lengths = BigInt64Array.from(list.map(c => BigInt(c.length)))
  const offsets = new BigInt64Array(list.length)
  let total = 0
  list.forEach((c, i) => { offsets[i] = BigInt(total); total += c.length })
  const samples = new Float32Array(total)
  list.forEach((c, i) => samples.set(c, Number(offsets[i])))
  return {samples, offsets, lengths}
}

function frameValue (out, N, i, feature) {
  if (feature === 'complexSpectrum') {
    return {real: out.complex_real.subarray(i * N, (i + 1) * N), imag: out.complex_imag.subarray(i * N, (i + 1) * N)}
  }
  if (feature === 'loudness') {
    const nb = out.numBarkBands || 24
    return {specific: out.loudness_specific.subarray(i * nb, (i + 1) * nb), total: out.loudness_total[i]}
  }
  const [field, len] = FIELD[feature]
  if (!len) return out[field][i]
  const n = len(N, out)
  return out[field].subarray(i * n, (i + 1) * n)
}

// extract(clips, {bufferSize, hop, sampleRate, windowingFunction, features, device, devices}, callback?)
function extract (clips, opts, callback) {
  const N = opts.bufferSize
  if (!isPowerOfTwo(N)) throw new Error('Buffer size is not a power of two: Meyda will not run.')  // src/meyda.js:20-22
  const features = splitFeatures(opts.features || FEATURES)
  const {samples, offsets, lengths} = packClips(clips)
  // opts.devices: clips sharded over several GPUs from this one process (mb_extract_multi; no inter-GPU traffic)
  if (Array.isArray(opts.devices) && opts.devices.length > 1) {
    const plans = opts.devices.map(d => planFor(Object.assign({}, opts, {device: d}), N, features))
    return wrapResult(native.extractMulti(plans, samples, offsets, lengths), N, features, callback)
  }
  const plan = planFor(opts, N, features)
  // opts.async: the blocking C-ABI call runs on the libuv pool (napi_async_work) and a Promise is returned
  if (opts.async) return native.extractAsync(plan, samples, offsets, lengths).then(out => wrapResult(out, N, features, callback))
  return wrapResult(native.extract(plan, samples, offsets, lengths), N, features, callback)
}

// extractAsync(clips, opts, callback?) -> Promise of the same result; the event loop stays free while the GPU works
function extractAsync (clips, opts, callback) { return extract(clips, Object.assign({}, opts, {async: true}), callback) }

// extractWav(files /* Uint8Array | Uint8Array[] of 16-bit PCM WAV files */, {bufferSize, hop, windowingFunction,
// features, channel, device}, callback?): BufferLoader + decodeAudioData + getChannelData(channel)
// (lib/bufferLoader.js:13-44, src/meyda.js:72) in front of the same pipeline; the int16 samples go to the GPU as
// they are and become s / 32768 inside the kernels' frame load.
function extractWav (files, opts, callback) {
  const N = opts.bufferSize
  if (!isPowerOfTwo(N)) throw new Error('Buffer size is not a power of two: Meyda will not run.')
  const features = splitFeatures(opts.features || FEATURES)
  const list = Array.isArray(files) ? files : [files]
  const infos = list.map(b => native.wavInfo(b))
  infos.forEach(i => {
    if (i.format !== 1 || i.bitsPerSample !== 16) throw new Error('extractWav takes 16-bit integer PCM')
    if (i.channels !== infos[0].channels || i.sampleRate !== infos[0].sampleRate) throw new Error('all WAV files of one call must share channel count and sample rate')
  })
  const ch = infos[0].channels
  const lengths = BigInt64Array.from(infos.map(i => BigInt(i.sampleFrames)))
  const offsets = new BigInt64Array(list.length)
  let total = 0  // clips start on multiples of 8 sample frames: mono frames stay 16-byte aligned for the TMA loads
  infos.forEach((i, c) => { offsets[c] = BigInt(total); total += Math.ceil(i.sampleFrames / 8) * 8 })
  const pcm = new Int16Array(total * ch)
  list.forEach((b, c) => pcm.set(new Int16Array(b.buffer, b.byteOffset + infos[c].dataOffset, infos[c].sampleFrames * ch), Number(offsets[c]) * ch))
  const plan = planFor(Object.assign({}, opts, {sampleRate: infos[0].sampleRate}), N, features)
  return wrapResult(native.extractPcm16(plan, pcm, ch, opts.channel || 0, offsets, lengths), N, features, callback)
}

// The reference's class (src/meyda.js:15-263) over the streaming entry points: `new Meyda(audioContext, src, bufSize,
// callback)`, get / start / stop / setSource / windowingFunction / featureInfo.  There is no Web Audio graph in Node:
// `audioContext` needs only `.sampleRate` (src/meyda.js:29), and a source is anything that hands over blocks of
// samples -- `src.connect(meyda)` (an EventEmitter-style source calls meyda.process(block)) or process(block) called
// directly, once per buffer, which is what onaudioprocess did (src/meyda.js:69-91).  Every block runs through
// mb_stream_push (one CUDA-graph replay per buffer, all 18 features); get() reads the current buffer's values.
class Meyda {
  constructor (audioContext, src, bufSize, callback) {
    if (!isPowerOfTwo(bufSize)) throw new Error('Buffer size is not a power of two: Meyda will not run.')  // src/meyda.js:20-22
    if (!audioContext) throw new Error("AudioContext wasn't specified: Meyda will not run.")               // src/meyda.js:24-26
    this.audioContext = audioContext
    this.bufferSize = bufSize || 256
    this.sampleRate = audioContext.sampleRate
    this.featureInfo = featureInfo
    this.EXTRACTION_STARTED = false
    this._featuresToExtract = null
    this._callback = callback
    this._window = 'hanning'  // src/meyda.js:41
    this._stream = null
    this._plan = null
    this._frame = null
    this.signal = null
    this.featureExtractors = Object.fromEntries(FEATURES.map(f => [f, {process: () => this.get(f)}]))
    if (src) this.setSource(src)
  }

  get windowingFunction () { return this._window }
  set windowingFunction (name) {  // a plan holds one window table: the next buffer starts a stream on the new one
    if (!(name in {hanning: 0, hamming: 1, blackman: 2})) throw new Error('unknown windowingFunction ' + name)
    if (name !== this._window) { this._window = name; this._dropStream() }
  }

  _dropStream () {
    if (this._stream) native.destroyStream(this._stream)
    this._stream = null
  }

  _ensureStream () {
    if (this._stream) return
    this._plan = planFor({sampleRate: this.sampleRate, windowingFunction: this._window}, this.bufferSize, FEATURES)
    this._stream = native.createStream(this._plan)
  }

  setSource (_src) {  // src/meyda.js:229-231
    if (_src && typeof _src.connect === 'function') _src.connect(this)
    if (this._stream) native.streamReset(this._stream)
  }

  start (features) { this._featuresToExtract = features; this.EXTRACTION_STARTED = true }  // src/meyda.js:233-236
  stop () { this._featuresToExtract = null; this.EXTRACTION_STARTED = false }              // src/meyda.js:238-241

  // One block of samples from the source (any length; the reference's ScriptProcessor hands over bufferSize at a
  // time).  Every buffer it completes becomes the current one in turn and fires the callback while started.
  process (block) {
    this._ensureStream()
    const out = native.streamPush(this._stream, this._plan, block instanceof Float32Array ? block : Float32Array.from(block))
    const n = Number(out.totalFrames)
    for (let i = 0; i < n; i++) {
      this._frame = {out, i}
      this.signal = out.buffer.subarray(i * this.bufferSize, (i + 1) * this.bufferSize)
      if (typeof this._callback === 'function' && this.EXTRACTION_STARTED) this._callback(this.get(this._featuresToExtract))
    }
    return n
  }

  get (feature) {  // src/meyda.js:244-261
    const value = f => {
      if (!featureInfo[f]) throw new TypeError('unknown feature ' + f)
      if (!this._frame) throw new Error('no buffer has been processed yet')
      return frameValue(this._frame.out, this.bufferSize, this._frame.i, f)
    }
    if (typeof feature === 'object' && feature !== null) {
      const results = {}
      for (let x = 0; x < feature.length; x++) {
        try { results[feature[x]] = value(feature[x]) } catch (e) { console.error(e) }
      }
      return results
    } else if (typeof feature === 'string') {
      return value(feature)
    }
    throw new Error('Invalid Feature Format')
  }
}

// Host threads a host-memory extract uses for the rows the device does not produce (`buffer`, powerSpectrum); 0 = automatic.
function setHostThreads(n) { native.setHostThreads(n | 0) }
// 2 / 1 / 0: those rows and the mirrored half of complexSpectrum, those rows only, or nothing on the host (the device
// produces the rest and it is copied back); -1 (default): 2 with one visible device and twelve or more cores, else 1.
function setHostRows(mode) { native.setHostRows(mode | 0) }
function getHostRows() { return native.getHostRows() }

module.exports = {Meyda, extract, extractAsync, extractWav, clearPlans, setHostThreads, setHostRows, getHostRows, featureInfo, isPowerOfTwo, FEATURES}
