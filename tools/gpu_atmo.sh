python - <<'PY'
import sys; sys.path.insert(0,"tests")
import numpy as np, shaderbox_b200 as sbx
from util import bits_equal
a=sbx.Renderer("APP_ATMOSPHERE",variant="plugin")
for v in ("plugin_k","plugin_k9"):
    b=sbx.Renderer("APP_ATMOSPHERE",variant=v)
    print(v, all(bits_equal(a.render(w,h,u_time=t), b.render(w,h,u_time=t)) for (w,h,t) in [(320,180,0.4),(200,111,1.0),(1920,1080,1.0)]))
PY
bash tools/bench_variants.sh atmosphere1080 plugin plugin_k plugin_k9
