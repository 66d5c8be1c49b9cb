for r in 1 2; do
echo "== main"; python tools/bench_configs.py c5 2>&1 | python tools/_fmt.py
for v in MB_BIG_LOAD_UNROLL=16 MB_BIG_LOAD_UNROLL=4; do echo "== $v"; MEYDA_B200_LIB=$PWD/meyda_b200/_lib/variants/lib_$v.so python tools/bench_configs.py c5 2>&1 | python tools/_fmt.py; done
done
