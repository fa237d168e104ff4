#!/bin/bash
# round 2, visit o (1 GPU): cooperative march with two sub-rounds per vote vs one
OUT=gpurun_out/r02o; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "coop or cooperative or uniform_space or far_from or default_variant" 2>&1 | tail -4 | tee $OUT/pytest.txt
timeout 600 python tools/part_time.py clouds1080 8 --variants coop,coop_u1 --splits rows4 2>&1 | tee $OUT/parts_n8.txt
timeout 600 python tools/part_time.py clouds1080 4 --variants coop2,coop2_u1 --splits rows4 2>&1 | tee $OUT/parts_n4.txt
python tools/variant_time.py clouds1080 coop coop_u1 coop2 coop2_u1 2>&1 | tee $OUT/variants.txt
echo done
