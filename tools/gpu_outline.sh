python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for w in vinyl1080 egg256 sdf_ao1080 atmosphere1080 planet2160 clouds1080 raytracer4320; do bash tools/bench_variants.sh $w plugin plugin_o 2>&1 | sed "s/^/$w /"; done
python - <<'PY'
import sys; sys.path.insert(0,"tests")
import numpy as np, shaderbox_b200 as sbx
from util import bits_equal
for app,w,h,t in [("APP_VINYL",320,180,1.25),("APP_EGG",200,120,2.75),("APP_SDF_AO",160,90,0.5),("APP_ATMOSPHERE",320,180,0.4),("APP_PLANET",256,144,5.5),("APP_CLOUDS",256,144,3.25),("APP_RAYTRACER",320,180,2.5)]:
    a=sbx.Renderer(app,variant="plugin").render(w,h,u_time=t); b=sbx.Renderer(app,variant="plugin_o").render(w,h,u_time=t)
    print(app, "outline == inline:", bits_equal(a,b))
PY
