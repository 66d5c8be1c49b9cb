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

// extract(clips, {bufferSize, hop, sampleRate, windowingFunction, features, device}, callback?)
function extract (clips, opts, callback) {
  const N = opts.bufferSize
  if (!isPowerOfTwo(N)) throw new Error('Buffer size is not a power of two: Meyda will not run.')  // src/meyda.js:20-22
  const features = splitFeatures(opts.features || FEATURES)
  const list = Array.isArray(clips) ? clips : [clips]
  const lengths = BigInt64Array.from(list.map(c => BigInt(c.length)))
  const offsets = new BigInt64Array(list.length)
  let total = 0
  list.forEach((c, i) => { offsets[i] = BigInt(total); total += c.length })
  const samples = new Float32Array(total)
  list.forEach((c, i) => samples.set(c, Number(offsets[i])))
  const featureMask = features.reduce((m, f) => m | (1 << FEATURES.indexOf(f)), 0)
  const plan = native.createPlan({
    bufferSize: N, hop: opts.hop || N, sampleRate: opts.sampleRate || 44100, featureMask,
    window: {hanning: 0, hamming: 1, blackman: 2}[opts.windowingFunction || 'hanning'] || 0, device: opts.device || 0,
    numBarkBands: opts.numBarkBands || 0, numMelFilters: opts.numMelFilters || 0,  // 0: the reference's 24 / 26 / 13 / 0.99
    numMfccCoefficients: opts.numMfccCoefficients || 0, rolloffFraction: opts.rolloffFraction || 0
  })
  const finish = out => {
    native.destroyPlan(plan)
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
  // opts.async: the blocking C-ABI call runs on the libuv pool (napi_async_work) and a Promise is returned
  if (opts.async) return native.extractAsync(plan, samples, offsets, lengths).then(finish, e => { native.destroyPlan(plan); throw e })
  return finish(native.extract(plan, samples, offsets, lengths))
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
  const featureMask = features.reduce((m, f) => m | (1 << FEATURES.indexOf(f)), 0)
  const plan = native.createPlan({
    bufferSize: N, hop: opts.hop || N, sampleRate: infos[0].sampleRate, featureMask,
    window: {hanning: 0, hamming: 1, blackman: 2}[opts.windowingFunction || 'hanning'] || 0, device: opts.device || 0,
    numBarkBands: opts.numBarkBands || 0, numMelFilters: opts.numMelFilters || 0,  // 0: the reference's 24 / 26 / 13 / 0.99
    numMfccCoefficients: opts.numMfccCoefficients || 0, rolloffFraction: opts.rolloffFraction || 0
  })
  const out = native.extractPcm16(plan, pcm, ch, opts.channel || 0, offsets, lengths)
  native.destroyPlan(plan)
  const result = {
    features, arrays: out, totalFrames: Number(out.totalFrames),
    value: (i, f) => frameValue(out, N, i, f),
    frame: i => Object.fromEntries(features.map(f => [f, frameValue(out, N, i, f)]))
  }
  if (typeof callback === 'function') for (let i = 0; i < result.totalFrames; i++) callback(result.frame(i))
  return result
}

module.exports = {extract, extractAsync, extractWav, featureInfo, isPowerOfTwo, FEATURES}
