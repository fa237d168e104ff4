"""Seeded inputs for the operator known-answer tests (sbx_eval_op / sbxref_eval_op layouts).

Each case: op name -> (in_stride, out_stride, generator(rng, n) -> float32 [n, in_stride]).
Ranges follow how the five apps call the operator (file:line of a typical call site)."""
import numpy as np


def _u(rng, n, k, lo, hi):
    return rng.uniform(lo, hi, size=(n, k)).astype(np.float32)


def _unit(rng, n):
    v = rng.normal(size=(n, 3)).astype(np.float32)
    v /= np.linalg.norm(v, axis=1, keepdims=True).astype(np.float32)
    return v.astype(np.float32)


def _cat(*a):
    return np.concatenate(a, axis=1).astype(np.float32)


def _special(rng, n):
    """floats incl. integers (lattice indices), tiny, huge, negative"""
    a = np.concatenate([
        rng.uniform(-4, 4, n // 4), rng.uniform(-120, 120, n // 4),
        np.round(rng.uniform(-3e5, 3e5, n // 4)), rng.uniform(-1e6, 1e6, n - 3 * (n // 4))]).astype(np.float32)
    return a[:, None]


CASES = {
    # libm layer (sbx_math.h)
    "sinf": (1, 1, _special),
    "cosf": (1, 1, _special),
    "tanf": (1, 1, lambda r, n: _u(r, n, 1, -20, 20)),
    "expf": (1, 1, lambda r, n: _u(r, n, 1, -100, 90)),
    "powf": (2, 1, lambda r, n: _cat(_u(r, n, 1, 0, 8), _u(r, n, 1, -4, 6))),
    "acosf": (1, 1, lambda r, n: _u(r, n, 1, -1.1, 1.1)),
    "atan2f": (2, 1, lambda r, n: _u(r, n, 2, -5, 5)),
    "sqrtf": (1, 1, lambda r, n: _u(r, n, 1, 0, 1e6)),
    "divf": (2, 1, lambda r, n: _u(r, n, 2, -100, 100)),
    # noise_iq.h:5-29, noise_worley.h:5-51, fbm.h:6,8
    "hash": (1, 1, lambda r, n: np.round(_u(r, n, 1, -2e5, 2e5))),
    "hash_arith": (1, 1, lambda r, n: _u(r, n, 1, -2e5, 2e5)),
    "noise_iq": (3, 1, lambda r, n: _u(r, n, 3, -300, 300)),
    "noise_w": (4, 3, lambda r, n: _cat(_u(r, n, 3, -8, 8), np.full((n, 1), 4.0))),
    "fbm4": (6, 1, lambda r, n: _cat(_u(r, n, 3, -40, 40), np.tile(np.float32([2.64, 0.5, 0.5]), (n, 1)))),
    "fbm_w3": (6, 1, lambda r, n: _cat(_u(r, n, 3, -4, 4), np.tile(np.float32([2.0, 0.5, 0.5]), (n, 1)))),
    # sdf.h:49-171, IK.h:44-52
    "sd_sphere": (4, 1, lambda r, n: _cat(_u(r, n, 3, -3, 3), _u(r, n, 1, 0.1, 2))),
    "sd_box": (6, 1, lambda r, n: _cat(_u(r, n, 3, -3, 3), _u(r, n, 3, 0.1, 2))),
    "sd_torus": (5, 1, lambda r, n: _cat(_u(r, n, 3, -3, 3), _u(r, n, 1, 0.5, 2), _u(r, n, 1, 0.05, 0.4))),
    "sd_y_cylinder": (5, 1, lambda r, n: _cat(_u(r, n, 3, -3, 3), _u(r, n, 2, 0.1, 2))),
    "sd_cylinder": (10, 1, lambda r, n: _cat(_u(r, n, 9, -3, 3), _u(r, n, 1, 0.05, 1))),
    "sd_bezier": (13, 2, lambda r, n: _cat(_u(r, n, 12, -3, 3), _u(r, n, 1, 0.02, 0.3))),
    "sd_capsule": (10, 1, lambda r, n: _cat(_u(r, n, 9, -3, 3), _u(r, n, 1, 0.05, 1))),
    "sd_plane": (7, 1, lambda r, n: _cat(_u(r, n, 3, -3, 3), _unit(r, n), _u(r, n, 1, -2, 2))),
    "op_blend": (3, 1, lambda r, n: _cat(_u(r, n, 2, -2, 2), _u(r, n, 1, 0.05, 1))),
    "ik_solver": (8, 3, lambda r, n: _cat(_u(r, n, 6, -2, 2), _u(r, n, 2, 0.3, 2))),
    # volumetric.h:5-45, util_optics.h:5-35, light.h:44-92, intersect.h:7-77
    "henyey_greenstein_phase_func": (1, 1, lambda r, n: _u(r, n, 1, -1, 1)),
    "rayleigh_phase_func": (1, 1, lambda r, n: _u(r, n, 1, -1, 1)),
    "schlick_phase_func": (1, 1, lambda r, n: _u(r, n, 1, -1, 1)),
    "isotropic_phase_func": (1, 1, lambda r, n: _u(r, n, 1, -1, 1)),
    "fresnel_factor": (3, 1, lambda r, n: _cat(_u(r, n, 2, 1, 2.5), _u(r, n, 1, 0, 1))),
    "reflect": (6, 3, lambda r, n: _cat(_unit(r, n), _unit(r, n))),
    "refract": (7, 3, lambda r, n: _cat(_unit(r, n), _unit(r, n), _u(r, n, 1, 0.4, 1.6))),
    "illum_cook_torrance": (14, 3, lambda r, n: _cat(_unit(r, n), _unit(r, n), _unit(r, n), _u(r, n, 3, 0, 1),
                                                       _u(r, n, 1, 0.05, 1), _u(r, n, 1, 1, 2.5))),
    "illum_blinn_phong": (14, 3, lambda r, n: _cat(_unit(r, n), _unit(r, n), _unit(r, n), _u(r, n, 3, 0, 1),
                                                     _u(r, n, 1, 0.05, 1), _u(r, n, 1, 1, 2.5))),
    "intersect_sphere": (10, 8, lambda r, n: _cat(_u(r, n, 3, -4, 4), _unit(r, n), _u(r, n, 3, -2, 2), _u(r, n, 1, 0.2, 3))),
    "intersect_plane": (10, 8, lambda r, n: _cat(_u(r, n, 3, -4, 4), _unit(r, n), _unit(r, n), _u(r, n, 1, -3, 3))),
    # util.h:5-138
    "rotate_around_x": (4, 3, lambda r, n: _cat(_u(r, n, 1, -720, 720), _u(r, n, 3, -3, 3))),
    "rotate_around_y": (4, 3, lambda r, n: _cat(_u(r, n, 1, -720, 720), _u(r, n, 3, -3, 3))),
    "rotate_around_z": (4, 3, lambda r, n: _cat(_u(r, n, 1, -720, 720), _u(r, n, 3, -3, 3))),
    "linear_to_srgb": (3, 3, lambda r, n: _u(r, n, 3, 0, 4)),
    "band": (4, 1, lambda r, n: _cat(_u(r, n, 2, 0, 1), _u(r, n, 2, 0.01, 0.3))),
    "checkboard_pattern": (3, 1, lambda r, n: _cat(_u(r, n, 2, -8, 8), _u(r, n, 1, 0.5, 4))),
    "remap": (5, 1, lambda r, n: _cat(_u(r, n, 1, -2, 2), _u(r, n, 4, -3, 3))),
    "get_primary_ray": (9, 6, lambda r, n: _cat(_u(r, n, 2, -1, 1), np.full((n, 1), -1.0), _u(r, n, 6, -5, 5))),
    "smoothstep": (3, 1, lambda r, n: _u(r, n, 3, -1, 2)),
    "mod": (2, 1, lambda r, n: _cat(_u(r, n, 1, -50, 50), _u(r, n, 1, 0.1, 7))),
    "fast_orthonormal_basis": (3, 6, lambda r, n: _unit(r, n)),
}


def inputs(op, n, seed=1234):
    in_stride, out_stride, gen = CASES[op]
    rng = np.random.default_rng(seed + sum(map(ord, op)))
    a = np.ascontiguousarray(gen(rng, n), dtype=np.float32)
    assert a.shape == (n, in_stride), (op, a.shape)
    return a, out_stride
