import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import shaderbox_b200 as sbx
from oracle import loader
from shaderbox_b200 import abi
a = loader.oracle_bake_volume(16); b = np.ascontiguousarray(a[::-1])
r = sbx.Renderer("APP_CLOUDS_TEX", variant=os.environ.get("V", "native"))
r.set_noise_volumes(a, b)
loader.oracle_set_noise_volumes(a, b)
w, h = int(os.environ.get("W", 96)), int(os.environ.get("H", 54))
rec = torch.zeros((4096, 8), dtype=torch.int64, device="cuda")
r.set_trace_buffer(rec.data_ptr())
got = r.render(w, h, u_time=1.5)
want = loader.oracle_render("APP_CLOUDS_TEX", abi.default_params(w, h, 1.5))
print("equal", np.array_equal(got.view(np.uint32), want.view(np.uint32)), float(np.nanmax(np.abs(got - want))))
rr = rec.cpu().numpy(); bad = rr[rr[:, 0] == 0xdead]
print("timeouts:", len(bad))
for x in bad[:12]:
    print("  wait parity %d mask %08x box (%d,%d,%d) state parity %d block %d thread %d" % (x[1], x[2], x[3] & 0xffffffff, x[3] >> 32, x[4], x[5], x[6], x[7]))
