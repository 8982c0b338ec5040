for v in base m24 m28; do python tools/variants.py run --players 8 --envs 4194304 --steps 512 $v; done
for v in base m28; do python tools/variants.py run --players 8 --envs 16777216 --steps 256 --preroll 1024 $v; done
