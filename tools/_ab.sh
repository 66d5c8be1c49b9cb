F=rms,energy,zcr,complexSpectrum,amplitudeSpectrum,spectralCentroid,spectralFlatness,spectralSlope,spectralRolloff,spectralSpread,spectralSkewness,spectralKurtosis,loudness,perceptualSpread,perceptualSharpness,mfcc
for f in all $F complexSpectrum rms,mfcc; do
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-secondary --features $f 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); e = d['e2e']
print('%-20s e2e %.3f M  ceiling %.3f M  frac %.3f  d2h %d MB h2d %d MB  d2h rate ceiling %.1f GB/s' % ('$f'[:20], e['value']/1e6, e['pcie_ceiling']['value']/1e6, e['frac_of_pcie_ceiling'], e['d2h_bytes_per_step']/1e6, e['h2d_bytes_per_step']/1e6, e['pcie_ceiling']['d2h_GBps_per_rank']))"
done
