for r in 1 2 4 8; do python tools/variants.py run --env SKYJO_RANGES=$r base; done
python tools/variants.py run --env SKYJO_RANGES=4 --players 8 --envs 4194304 base
python tools/variants.py run --env SKYJO_RANGES=1 --players 8 --envs 4194304 base
python tools/variants.py run --env SKYJO_RANGES=4 --players 2 base
python tools/variants.py run --env SKYJO_RANGES=1 --players 2 base
