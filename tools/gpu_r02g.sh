#!/bin/bash
# round 2, visit g (1 GPU): floor-without-FRND experiment; native CLOUDS ncu with source counters (lanes per instruction)
OUT=gpurun_out/r02g; mkdir -p $OUT
timeout 300 python - <<'PY' 2>&1 | tee $OUT/fm_parity.txt
import numpy as np, shaderbox_b200 as sbx
from tests.util import bits_equal
a = sbx.Renderer("APP_CLOUDS", variant="native"); b = sbx.Renderer("APP_CLOUDS", variant="native_fm")
for (w, h, t, ov) in ((1920, 1080, 1.5, {"cld_march_steps": 128}), (333, 187, 77.0, {"wind_dir": (0.1, 0.0, 0.3)}), (256, 144, 5000.0, {})):
    x, y = a.render(w, h, u_time=t, **ov), b.render(w, h, u_time=t, **ov)
    print("floor-magic vs FRND %dx%d t=%g: bits equal %s, max|d| %g" % (w, h, t, bits_equal(x, y), float(np.nanmax(np.abs(x - y)))))
PY
python tools/variant_time.py clouds1080 native native_fm native native_fm 2>&1 | tee $OUT/variants.txt
for v in native native_fm; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sbx_render -s 3 -c 1 -o $OUT/prof_clouds1080_$v \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --variant $v > $OUT/ncu_$v.log 2>&1; tail -1 $OUT/ncu_$v.log
done
echo done
