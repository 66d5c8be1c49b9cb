python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
echo "== main (lock 4)"; python tools/bench_configs.py smallab c3 c1 2>&1 | python tools/_fmt.py
for v in MB_NO_NOISE_STATS; do echo "== $v"; MEYDA_B200_LIB=$PWD/meyda_b200/_lib/variants/lib_$v.so python tools/bench_configs.py smallab c3 c1 2>&1 | python tools/_fmt.py; done
