"""BASELINE config 5 (projection on a fixed grid with walls and sources) on the
CPU: the synthetic input is reproducible window by window, and the oracle pins the residual target the
GPU solver is held to (max|div| of the red-black solves <= the reference's lexicographic sweeps)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_splitmix_stream_known_answers_and_windows():
    from fluid_b200 import presets
    # SplitMix64(seed = 0x5EED): first outputs computed by hand from the published algorithm
    def ref(k, seed=0x5EED):
        m = (1 << 64) - 1
        z = (seed + (k + 1) * 0x9E3779B97F4A7C15) & m
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
        z ^= z >> 31
        return np.float32((z >> 40) / float(1 << 24) * 2.0 - 1.0)
    a = presets.splitmix_uniform(64, 0x5EED)
    assert all(a[k] == ref(k) for k in range(64))
    assert a.min() >= -1.0 and a.max() < 1.0
    b = presets.splitmix_uniform(20, 0x5EED, start=37)
    assert np.array_equal(a[37:57], b)
    u, v = presets.projection_fields(12, 9, 4, 5)
    full = presets.splitmix_uniform(2 * 12 * 9, 0x5EED)
    assert np.array_equal(u, full[:108].reshape(12, 9)[4:9])
    assert np.array_equal(v, full[108:].reshape(12, 9)[4:9])


def test_torch_generator_equals_numpy_generator():
    import torch

    import bench
    from fluid_b200 import presets
    a = bench.splitmix_uniform_torch(torch, 4096, 0x5EED, 32770 * 32770 - 100, "cpu").numpy()   # offsets beyond 2^30 too
    assert np.array_equal(a, presets.splitmix_uniform(4096, 0x5EED, 32770 * 32770 - 100))


def test_config5_residual_target_on_the_oracle():
    """Lexicographic 8 sweeps (the reference) against the red-black solves on the config-5 input: the
    max|div| the GPU path must reach, and the pressure form equals the face form within rounding."""
    import oracle
    from fluid_b200 import presets
    size = (192, 160)
    p = presets.projection_stress(*size)
    u, v = presets.projection_fields(size[0] + 2, size[1] + 2, 0, size[0] + 2)
    res = {}
    fields = {}
    for name, solver in (("lex", oracle.SOLVER_EXACT), ("rb", oracle.SOLVER_REDBLACK), ("rbq", oracle.SOLVER_REDBLACK_PRESSURE)):
        f = oracle.New(p.density, p.width, p.height, p.h, solver=solver)
        f.set("U", u); f.set("V", v)
        f.edit(p.init); f.edit(p.per_step)
        before = f.MaxDivergence()
        # Q-16: faces of solid cells are zero after SetSolid(true)
        S = f.get("S")
        assert np.all(f.get("U")[S == 0] == 0) and np.all(f.get("V")[S == 0] == 0)
        f.project(8, p.dt)
        res[name] = (before, f.MaxDivergence())
        fields[name] = (f.get("U"), f.get("V"))
    assert res["lex"][0] == res["rb"][0] == res["rbq"][0] > 1.0
    assert res["rb"][1] <= res["lex"][1] < res["lex"][0]
    assert res["rbq"][1] <= res["lex"][1]
    for a, b in zip(fields["rb"], fields["rbq"]):
        assert np.max(np.abs(a - b)) <= 2e-5


def test_bench_reference_arm_of_the_projection_workload():
    import bench
    r = bench.projection_cpu_run(128, 1, 0)
    assert r["cores"] == 1 and r["value"] > 0
    assert r["max_div_after_8_sweeps"] < r["max_div_before"]
