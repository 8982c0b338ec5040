for d in 0 1184 2368 4736 9472; do python tools/variants.py run --env SKYJO_PF_DIST=$d base; done
for v in keep stream keepstream w24 w28; do python tools/variants.py run --env SKYJO_PF_DIST=0 $v; python tools/variants.py run --env SKYJO_PF_DIST=4736 $v; done
