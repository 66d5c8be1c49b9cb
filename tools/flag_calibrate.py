"""Calibration of the adaptive-exactness criteria of meyda_b200/csrc/mb_adaptive.cuh (numpy prototype, CPU only).

Emulates the float32-FFT path (scipy float32 FFT + the oracle's epilogue) against the reference arithmetic (oracle
jsfft) on the demo clips, the bench's synthetic clips, pure tones with and without a noise floor, low-passed noise and
its 16-bit quantisation; per feature it prints v = frames outside the flat 1e-3 tolerance, f = frames the criterion
flags, miss = violating frames not flagged by that criterion / by any criterion.  The criteria must show miss 0
everywhere while leaving ordinary audio (sound2, the synthetic clips, quantised audio) unflagged.
Usage: python tools/flag_calibrate.py [K=16 theta=32 kap=8 ...]"""
import sys, numpy as np
sys.path.insert(0, '/root/repo')
from oracle import meyda_oracle as mo
import scipy.fft as sf
f64 = np.float64
SR = 44100.0
FEATS = ["spectralCentroid","spectralFlatness","spectralSlope","spectralSpread","spectralSkewness","spectralKurtosis","loudness","perceptualSpread","perceptualSharpness","mfcc"]

def fast_spec(windowed):
    z = sf.fft(windowed.astype(np.float32), axis=1)
    N = windowed.shape[1]
    z = np.conj(z) * np.float32(1/np.sqrt(N))
    return z.real.astype(np.float32), z.imag.astype(np.float32)

def violations(g, r, tol=1e-3, slope=False):
    g = np.asarray(g, f64); r = np.asarray(r, f64)
    fin = np.isfinite(r) & np.isfinite(g)
    mism = (np.isfinite(r) != np.isfinite(g)) | (~np.isfinite(r) & ~((np.isnan(g)&np.isnan(r)) | (g==r)))
    err = np.abs(np.where(fin, g-r, 0))
    ok = (err <= tol*np.abs(np.where(fin, r, 1))) | (err <= (1e-13 if slope else tol))
    return (~ok) | mism

def flags(amp, Eraw, N, K=16.0, theta=32.0, kap=8.0, m=0.5, t=1e-3, kb=2.0, k2=4.0, cf1=1.0, cb=1.0, kr=4.0, detail=False):
    """amp: [F, n] f64 fast amplitudes; returns dict of per-criterion flags"""
    F, n = amp.shape
    srms = 1e-7*np.sqrt(Eraw/N)          # rms-level estimate of |Zfast - Zref| per bin (x1.63 conservative: raw energy)
    sp = K*srms                          # hard per-bin bound
    out = {}
    k = np.arange(n, dtype=f64)
    # ---- loudness
    bb = mo.bark_band_limits(mo.bark_scale(N, SR), n)
    nbins = np.diff(bb).astype(f64)
    B = np.stack([amp[:, bb[b]:bb[b+1]].sum(1) for b in range(24)], 1)
    eB = (kb*nbins[None,:] + K*np.sqrt(nbins[None,:]))*srms[:,None]
    up = np.power(B+eB, 0.23); lo = np.power(np.maximum(B-eB, 0), 0.23)
    spec = np.power(B, 0.23)
    dspec = np.where(nbins[None,:] > 0, up-lo, 0.0)
    total = spec.sum(1); dtot = dspec.sum(1)
    f_spec = (dspec > m*t*np.maximum(1, spec)).any(1)
    f_tot = dtot > m*t*np.maximum(1, total)
    mx = spec.max(1); r = (total-mx)/total
    dr = (dspec.max(1) + dtot)/total
    f_psp = 2*r*dr > m*t
    w = np.zeros(24); w[1:16] = np.arange(1,16)
    sharp = 0.11*((spec*w).sum(1) + 19.9796966)/total
    dsh = 0.11*(dspec*w).sum(1)/total + sharp*dtot/total
    f_psh = dsh > m*t*np.maximum(1, sharp)
    out['loudness'] = f_spec | f_tot; out['perceptualSpread'] = f_psp; out['perceptualSharpness'] = f_psh
    # ---- mfcc
    fb = mo.mel_filterbank(N, SR)[:, :n].astype(f64)
    W = fb.sum(1)
    Ef = (amp**2) @ fb.T
    with np.errstate(all='ignore'):
        rf = srms[:,None]*np.sqrt(W[None,:]/Ef)
        d = 2*K*rf/np.sqrt(np.maximum(W[None,:],1)) + k2*rf*rf
        dl = np.where(W[None,:] > 0, np.where(d < 0.5, 2*d, np.inf), 0.0)
        dl = np.where(np.isnan(dl), np.inf, dl)
    f_mfcc = (0.2774/13)*dl.sum(1) > m*t
    out['mfcc'] = f_mfcc
    # ---- moments
    S = [ (amp*k**q).sum(1) for q in range(5) ]
    T = [ (k**q).sum() for q in range(9) ]
    with np.errstate(all='ignore'):
        qb = np.minimum((theta*srms[:,None])**2/(amp*amp), 1.0)       # min(1, (theta sigma / a)^2)
    qb = np.where(np.isnan(qb), 1.0, qb)
    nblk = max(n//32, 1); bs = n//nblk
    Qblk = qb.reshape(F, nblk, bs).sum(2)
    kmax = (np.arange(nblk)*bs + bs - 1).astype(f64)
    Q = [ (Qblk*kmax[None,:]**q).sum(1) for q in range(5) ]
    # bias: a floor bin moves by ~sigma, a bin at x sigma by sigma/x^2-ish: sigma * Q-weighted; random: kap sigma sqrt(T2q)
    dS = [ kap*srms*np.sqrt(T[2*q]) + cb*srms*Q[q] for q in range(5) ]
    with np.errstate(all='ignore'):
        # subsampled Q (8 of every 32 bins, irregular offsets), exponent-trick error folded as x1.6
        sel = np.zeros(n, bool)
        offs = [0, 7, 10, 13, 16, 23, 26, 29]
        for o in offs: sel[o::32] = True
        if n < 32: sel[:] = True
        Qblk_s = (qb*sel[None,:]).reshape(F, nblk, bs).sum(2) * (bs/ max(1, sel[:bs].sum())) * 1.6
        q0 = Qblk_s.sum(1); q4 = (Qblk_s*kmax[None,:]**4).sum(1)
        gq = np.where((q0>0)&(q4>0), np.sqrt(np.sqrt(q4/q0)), 0.0)
        inv0 = 1.0/S[0]
        sT = [kap*np.sqrt(T[2*q]) for q in range(5)]
        r0 = srms*(sT[0]+q0)*inv0
        q1 = q0*gq; q2 = q1*gq; q3 = q2*gq
        m1, m2, m3, m4 = (S[i]*inv0 for i in (1,2,3,4))
        d1 = srms*(sT[1]+q1)*inv0 + m1*r0
        d2 = srms*(sT[2]+q2)*inv0 + m2*r0
        d3 = srms*(sT[3]+q3)*inv0 + m3*r0
        d4 = srms*(sT[4]+q4)*inv0 + m4*r0
        var = m2-m1*m1; sd = np.sqrt(var); iv = 1/var; isd = 1/sd
        tl = m*t
        ok0 = S[0] > 0
        out['spectralCentroid'] = ~(d1 <= tl*np.maximum(1, m1)) & ok0
        out['spectralSlope'] = ~(d1 <= tl*np.abs(m1-0.5*(n-1))) & ok0
        dv = d2 + 2*m1*d1
        badv = ~(dv <= 0.25*var)
        out['spectralSpread'] = (badv | ~(0.5*dv*isd <= tl*np.maximum(1, sd))) & ok0
        A = 2*m1**3 - 3*m1*m2 + m3; c = 1.5*A*iv
        e = (np.abs(6*m1*m1-3*m2+c*2*m1)*d1 + np.abs(-3*m1-c)*d2 + d3)*iv*isd
        out['spectralSkewness'] = (badv | ~(e <= tl*np.maximum(1, np.abs(A*iv*isd)))) & ok0
        B = -3*m1**4 + 6*m1*m2 - 4*m1*m3 + m4; c = 2*B*iv
        e = (np.abs(-12*m1**3+6*m2-4*m3+c*2*m1)*d1 + np.abs(6*m1-c)*d2 + 4*m1*d3 + d4)*iv*iv
        out['spectralKurtosis'] = (badv | ~(e <= tl*np.maximum(1, np.abs(B*iv*iv)))) & ok0
        L = np.log(amp).sum(1)
        flat = np.exp(L/n)*n/S[0]
        dml = (q0 + 4*np.sqrt(q0)/theta)/n + r0
        out['spectralFlatness'] = (~(dml <= 1.0) | ~((dml+dml*dml) <= tl) | (np.isneginf(L))) & ok0
    if detail: out['_dd'] = dd
    return out

def run(x, N, hop, label, **kw):
    fr = mo.frame_signal(x, N, hop)
    if len(fr) > 1500: fr = fr[::len(fr)//1500]
    win = mo.window_table(N, 'hanning')
    w = (fr.astype(f64)*win).astype(np.float32)
    ref = mo.extract_frames(fr, SR, 'hanning', FEATS)
    fs = fast_spec(w)
    fast = mo.extract_frames(fr, SR, 'hanning', FEATS, fs)
    n = N//2
    amp = np.sqrt(fs[0][:, :n].astype(f64)**2 + fs[1][:, :n].astype(f64)**2).astype(np.float32).astype(f64)
    Eraw = (fr.astype(f64)**2).sum(1)
    fl = flags(amp, Eraw, N, **kw)
    anyflag = np.zeros(len(fr), bool)
    for kk in FEATS: anyflag |= fl[kk]
    res = []
    for kk in FEATS:
        if kk == 'loudness':
            v = violations(fast[kk]['specific'], ref[kk]['specific']).any(1) | violations(fast[kk]['total'], ref[kk]['total'])
        elif kk == 'mfcc':
            v = violations(fast[kk], ref[kk]).any(1)
        else:
            v = violations(fast[kk], ref[kk], slope=(kk=='spectralSlope'))
        res.append('%s v%d f%d miss%d/%d' % (kk[:12], v.sum(), fl[kk].sum(), (v & ~fl[kk]).sum(), (v & ~anyflag).sum()))
    print('%-28s N=%5d frames=%5d anyflag=%5d (%.1f%%) | ' % (label, N, len(fr), anyflag.sum(), 100*anyflag.mean()) + ' ; '.join(res))

if __name__ == '__main__':
    g = np.load('/root/repo/tests/golden/audio_pcm16.npz')
    kw = {}
    for a in sys.argv[1:]:
        k_, v_ = a.split('='); kw[k_] = float(v_)
    for name in ['sound1','sound2','sound3']:
        x = mo.pcm16_to_float(g[name])
        for N in [256, 512, 1024, 2048]:
            run(x, N, N, name, **kw)
    for i in range(3):
        run(mo.synth_clip(i, 2048*40), 2048, 512, 'synth%d' % i, **kw)
    rng = np.random.default_rng(1)
    t = np.arange(2048*60)/SR
    for f0, nz in [(440.0, 0), (440.0, 1e-6), (440.0, 1e-4), (3000.3, 1e-5), (12000.7, 0), (50.0, 1e-3)]:
        x = (0.5*np.sin(2*np.pi*f0*t) + nz*rng.standard_normal(len(t))).astype(np.float32)
        run(x, 2048, 512, 'tone%g+nz%g' % (f0, nz), **kw)
        run(x, 512, 512, 'tone%g+nz%g' % (f0, nz), **kw)
    # low-passed noise (like mp3-sourced audio)
    z = np.fft.rfft(rng.standard_normal(2048*60)); z[len(z)*16//22:] = 0
    x = (0.1*np.fft.irfft(z)/30).astype(np.float32)
    run(x, 2048, 512, 'lowpass16k', **kw)
    x16 = np.round(x*32768).astype(np.int16)
    run(mo.pcm16_to_float(x16), 2048, 512, 'lowpass16k-pcm16', **kw)
