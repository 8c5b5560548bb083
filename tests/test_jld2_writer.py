"""The result-file writer (piql.jl_b200/jld2.py, `@save "data/<name>.jld2" out` of scripts/distributed_experiments.jl:49-59)
against the decoder that reads the reference's own data/*.jld2 (tools/make_golden.py): a result tuple rebuilt from the golden
fixtures, written and decoded again, must give the fixtures back bit for bit; the container carries valid lookup3 checksums."""
import importlib.util
import os
import struct

import numpy as np

from conftest import CONFIGS, ROOT, Golden


def _decoder():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tools", "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _result(piql, g):
    """The dict solvers.run_experiment returns, filled with the reference's stored values."""
    gw = piql.Gridworld(np.asfortranarray(g.walls != 0), g.goal)
    ctrl = piql.gridworld_ctrl(gw, randomness=g.params["randomness"], gamma=g.params["gamma"], step_cost=g.params["step_cost"],
                               reward=g.params["reward"])
    G = lambda a: piql.Growable.from_dense(a, g.walls)
    betas = [float(b) for b in g.betas]
    return dict(nu=G(g.nu), traj={b: [G(g.V(i, k)) for k in range(6)] for i, b in enumerate(betas)},
                gap={b: [G(g.G(i, k)) for k in range(6)] for i, b in enumerate(betas)},
                randomness=g.params["randomness"], gamma=g.params["gamma"], ctrl=ctrl, s0=g.s0)


def test_written_results_decode_to_the_reference_values(piql, tmp_path):
    dec = _decoder()
    for name in CONFIGS:
        g = Golden(name)
        path = piql.jld2.save_result(str(tmp_path / (name + ".jld2")), _result(piql, g))
        d = dec.decode(path)
        ref = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        for key in ("walls", "goal", "s0", "betas", "states", "nu", "traj", "gap"):
            assert np.array_equal(d[key], ref[key]), (name, key)
        for key in ("randomness", "gamma", "step_cost", "reward"):
            assert float(d[key]) == float(ref[key]), (name, key)


def test_container_structure_and_checksums(piql, tmp_path):
    g = Golden("r0_01g0_99")
    path = piql.jld2.save_result(str(tmp_path / "x.jld2"), _result(piql, g))
    b = open(path, "rb").read()
    assert b.startswith(b"HDF5-based Julia Data Format, version 0.1.1") and b[512:520] == b"\x89HDF\r\n\x1a\n"
    sb = b[512:512 + 48]
    assert sb[8] == 2 and sb[9] == 8 and sb[10] == 8
    base, ext, eof, root = struct.unpack("<4Q", sb[12:44])
    assert base == 512 and eof == len(b) and struct.unpack("<I", sb[44:48])[0] == piql.jld2.lookup3(sb[:44])
    # every object header: "OHDR", version 2, and the lookup3 checksum behind its message block
    n, p = 0, b.find(b"OHDR", 512)
    while p >= 0:
        assert b[p + 4] == 2 and b[p + 5] == 2
        size = struct.unpack("<I", b[p + 6:p + 10])[0]
        end = p + 10 + size
        assert struct.unpack("<I", b[end:end + 4])[0] == piql.jld2.lookup3(b[p:end])
        n, p = n + 1, b.find(b"OHDR", end)
    assert n > 200        # 1 + 48 + 48 Growables (two objects each), the two dictionaries, walls, goal, out, the root group
    # lookup3 known answers (Bob Jenkins' lookup3.c self-test: hashlittle of the empty string and of "Four score ...")
    assert piql.jld2.lookup3(b"") == 0xDEADBEEF
    assert piql.jld2.lookup3(b"Four score and seven years ago", 0) == 0x17770551
    assert piql.jld2.lookup3(b"Four score and seven years ago", 1) == 0xCD628161
