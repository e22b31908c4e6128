"""Static checks on the SASS of the built library (cuobjdump, no GPU needed): the optimistic and the
guarded form of every OPTIMISTIC pair loop must do the same FP32 work, and the work must be the
LANE_OPS / SFU_OPS the roofline is computed from."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.fixture(scope="module")
def loops():
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    import sass_mix
    lib = os.path.join(ROOT, "cvortex_b200", "lib", "libcvortex.so")
    if not os.path.exists(lib):
        from cvortex_b200 import _native
        _native.build(jobs=8)
    text = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    return sass_mix.loop_table(text)


def test_plain_and_guarded_loops_do_the_same_fp32_work(loops):
    """nvcc contracts a packed product into a following packed sum when nothing stands between them
    -- which a guard's select does and its absence does not.  A loop pair whose forms differ in
    lane-ops therefore rounds differently (F3D dvort did, before its B sum became one explicit FMA)."""
    pairs = {}
    for row in loops:
        pairs.setdefault((row["policy"], row["T"]), {})[row["form"]] = row
    both = {k: v for k, v in pairs.items() if len(v) == 2 and not k[0].startswith("F3D")}
    assert len(both) == 4 * 8, sorted(both)           # 8 optimistic policies x 4 geometries
    for key, v in both.items():
        assert v["plain"]["lane_ops"] == v["guarded"]["lane_ops"], (key, v)
        assert v["plain"]["mufu"] == v["guarded"]["mufu"], (key, v)
        assert v["plain"]["alu"] == 0, (key, v)


def test_loops_match_the_declared_work_per_pair(loops):
    from cvortex_b200 import api
    lib = api.library()
    ops = {"P3DVel": "P3D_M2M_vel", "P3DDvort": "P3D_M2M_dvort", "P3DVisc": "P3D_M2M_visc_dvort", "P3DVort": "P3D_M2M_vort",
           "P2DVel": "P2D_M2M_vel", "P2DVisc": "P2D_M2M_visc_dvort", "F3DVel": "F3D_M2M_vel", "F3DDvort": "F3D_M2M_dvort",
           "P3DVelDvort": "P3D_M2M_vel_dvort"}
    regs = ["singular", "winckelmans", "planetary", "gaussian"]
    seen = 0
    for row in loops:
        name, _, reg = row["policy"].partition("<")
        reg = regs[int(reg[:-1])] if reg else "singular"
        info = api.backend().op_info(ops[name], reg)
        if name == "P3DVort" and reg == "singular":
            continue                                  # zeta = 0: the compiler deletes the loop body
        if name.startswith("F3D"):
            # filament tiers: the declared work is the cancellation-free form's; the loop also carries one scalar
            # multiply per SOURCE (tau l^2), i.e. 1/T per pair.  The loop that selects per pair (F3D_WIDE): the
            # reference's rounded cross product, 39 / 55 lane-ops, 4 / 5 MUFU, a compare + select next to the flag's minimum.
            wide = {"F3DVel": (39, 4), "F3DDvort": (55, 5)}[name]
            want = {"new": (info["lane_ops"], info["sfu_ops"], 1.0), "wide": (wide[0], wide[1], 2.0)}[row["form"]]
            assert want[0] <= row["lane_ops"] <= want[0] + 1.0 / row["T"] + 1e-9, (row, info)
            assert row["mufu"] == want[1] and row["alu"] == want[2], (row, info)
            seen += 1
            continue
        assert row["lane_ops"] == info["lane_ops"], (row, info)
        assert row["mufu"] == info["sfu_ops"], (row, info)
        seen += 1
    assert seen >= 100 and lib is not None
