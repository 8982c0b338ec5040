(cd build/old_tree && python tools/quick_bench.py --players 8 --envs 16777216 --steps 64 --tag old16M; python tools/quick_bench.py --players 8 --envs 4194304 --steps 128 --tag old4M; python tools/quick_bench.py --players 8 --envs 4194304 --steps 512 --tag old4M-512)
python tools/quick_bench.py --players 8 --envs 4194304 --steps 512 --tag new4M-512
SKYJO_PF_DIST=0 python tools/quick_bench.py --players 8 --envs 4194304 --steps 512 --tag new4M-512-pf0
