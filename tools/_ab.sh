for r in 1 2; do
echo "== main (lock 0)"; python tools/bench_configs.py c3 pcm 2>&1 | python tools/_fmt.py
for v in MB_LOCK_WARPS=2 MB_LOCK_WARPS=4 MB_LOCK_WARPS=8; do echo "== $v"; MEYDA_B200_LIB=$PWD/meyda_b200/_lib/variants/lib_$v.so python tools/bench_configs.py c3 pcm 2>&1 | python tools/_fmt.py; done
done
