#!/usr/bin/env python3
"""Decode the reference's result files (data/*.jld2) into flat golden fixtures.

Run in the build container only (it reads /root/reference, which does not exist
on the GPU box); the fixtures it writes under tests/golden/ are committed.

    python tools/make_golden.py [/root/reference/data] [tests/golden]

The files were written by scripts/distributed_experiments.jl:49-59 of the
reference (JLD2 format 0.1.1, Julia 1.10.0): one NamedTuple `out` =
(ν, traj, gap, randomness, γ, ctrl, s0) per (randomness, γ) configuration.
JLD2 is an HDF5 subset: a 512-byte text header, then a version-2 superblock and
version-2 object headers whose datasets here are all stored with the *compact*
layout.  No HDF5 library is needed (none is installed): the ~150 lines below
walk the object headers directly.  Every stored address is relative to the
superblock base (file offset = address + 512).

Output per file, `tests/golden/<name>.npz`:
    walls      uint8 (H, W)   1 = wall, indexed [i-1, j-1] like the Julia array
    goal, s0   int64 (2,)     1-based (i, j) as in Julia
    randomness, gamma, step_cost, reward   float64 scalars
    betas      float64 (8,)   ascending
    states     int64 (251, 2) 1-based (i, j), sorted column-major
    nu         float64 (251,)          generate_ν(ctrl)
    traj       float64 (8, 6, 251)     traj[b, k] = V_k for betas[b]; k=0 is V_lo
    gap        float64 (8, 6, 251)     generate_optimality_gap of traj[b, k]
"""
import os
import struct
import sys

import numpy as np

BASE = 512


class Jld2:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        sb = self.b[BASE:]
        assert sb[:8] == b"\x89HDF\r\n\x1a\n" and sb[8] == 2, "not a v2 superblock"
        assert sb[9] == 8 and sb[10] == 8
        _base, _ext, _eof, self.root = struct.unpack("<4Q", sb[12:44])

    # -- object header v2 ---------------------------------------------------
    def messages(self, addr):
        """Yield (type, flags, payload) for every message of the object at addr."""
        b = self.b
        o = addr + BASE
        assert b[o:o + 4] == b"OHDR", (hex(addr), b[o:o + 4])
        flags = b[o + 5]
        p = o + 6
        if flags & 0x20:
            p += 16
        if flags & 0x10:
            p += 4
        szlen = 1 << (flags & 3)
        chunk = int.from_bytes(b[p:p + szlen], "little")
        p += szlen
        blocks = [(p, p + chunk)]
        has_order = bool(flags & 0x04)
        while blocks:
            p, end = blocks.pop(0)
            while p + 4 <= end:
                mtype = b[p]
                msize = struct.unpack("<H", b[p + 1:p + 3])[0]
                mflags = b[p + 3]
                p += 4 + (2 if has_order else 0)
                data = b[p:p + msize]
                p += msize
                if mtype == 0x10:  # continuation -> OCHK block
                    caddr, clen = struct.unpack("<QQ", data[:16])
                    co = caddr + BASE
                    assert b[co:co + 4] == b"OCHK"
                    blocks.append((co + 4, co + clen - 4))
                elif mtype != 0:
                    yield mtype, mflags, data

    def dataset(self, addr):
        """Return (dims, raw bytes) of the dataset at addr."""
        dims, raw = (), None
        for mtype, _mf, d in self.messages(addr):
            if mtype == 0x01:  # dataspace v2
                assert d[0] == 2
                rank = d[1]
                dims = struct.unpack("<%dQ" % rank, d[4:4 + 8 * rank])
            elif mtype == 0x08:  # layout v4 (v3 also seen in JLD2)
                ver, cls = d[0], d[1]
                assert ver in (3, 4), ver
                if cls == 0:  # compact
                    n = struct.unpack("<H", d[2:4])[0]
                    raw = d[4:4 + n]
                elif cls == 1:  # contiguous
                    a, n = struct.unpack("<QQ", d[2:18])
                    raw = self.b[a + BASE:a + BASE + n]
                else:
                    raise ValueError("chunked layout not expected")
        return dims, raw

    def links(self, addr):
        out = {}
        for mtype, _mf, d in self.messages(addr):
            if mtype == 0x06:  # link message v1, hard link
                lflags = d[1]
                p = 2
                if lflags & 0x08:
                    p += 1
                if lflags & 0x04:
                    p += 8
                if lflags & 0x10:
                    p += 1
                nlen_sz = 1 << (lflags & 3)
                nlen = int.from_bytes(d[p:p + nlen_sz], "little")
                p += nlen_sz
                name = d[p:p + nlen].decode()
                p += nlen
                out[name] = struct.unpack("<Q", d[p:p + 8])[0]
        return out

    # -- the structures this reference stored --------------------------------
    def growable(self, ref):
        """Follow 8-byte refs until an array of {i::Int64, j::Int64, v::Float64}."""
        for _ in range(6):
            dims, raw = self.dataset(ref)
            if raw is not None and len(dims) == 1 and len(raw) == 24 * dims[0] and dims[0] > 1:
                rec = np.frombuffer(raw, dtype=np.dtype([("i", "<i8"), ("j", "<i8"), ("v", "<f8")]))
                return rec
            assert raw is not None and len(raw) >= 8, (hex(ref), dims)
            ref = struct.unpack("<Q", raw[:8])[0]
        raise ValueError("no Growable found under %x" % ref)

    def beta_dict(self, ref):
        """Dict{Float64, Vector{Growable}} -> {beta: [records]*6}."""
        # ref -> (8-byte ref ->)* array(8) of refs -> {beta, ref} -> array(6) of refs
        for _ in range(6):
            dims, raw = self.dataset(ref)
            if len(dims) == 1 and dims[0] == 8 and len(raw) == 64:
                break
            ref = struct.unpack("<Q", raw[:8])[0]
        else:
            raise ValueError("beta dict not found")
        out = {}
        for pref in struct.unpack("<8Q", raw):
            _d, praw = self.dataset(pref)
            beta, vref = struct.unpack("<dQ", praw[:16])
            for _ in range(6):
                vdims, vraw = self.dataset(vref)
                if len(vdims) == 1 and vdims[0] == 6 and len(vraw) == 48:
                    break
                vref = struct.unpack("<Q", vraw[:8])[0]
            else:
                raise ValueError("iterate vector not found")
            out[beta] = [self.growable(r) for r in struct.unpack("<6Q", vraw)]
        return out


def decode(path):
    f = Jld2(path)
    out_addr = f.links(f.root)["out"]
    dims, raw = f.dataset(out_addr)
    assert len(raw) == 216, len(raw)
    w = struct.unpack("<27Q", raw)
    fl = struct.unpack("<27d", raw)
    randomness, gamma = fl[3], fl[4]
    step_cost, reward = fl[10], fl[11]
    wd, wraw = f.dataset(w[9])
    assert len(wd) == 2 and len(wraw) == wd[0] * wd[1]
    # HDF5 dims are listed slowest-first; JLD2 reverses Julia's dims, and the bytes are
    # Julia column-major: walls[i, j] at (i-1) + H*(j-1).
    W_, H_ = wd
    walls = np.frombuffer(wraw, dtype=np.uint8).reshape(W_, H_).T.copy()
    _gd, graw = f.dataset(w[13])
    goal = np.array(struct.unpack("<2q", graw[:16]), dtype=np.int64)
    s0 = np.array(struct.unpack("<2q", raw[25 * 8:27 * 8]), dtype=np.int64)

    nu = f.growable(w[0])
    traj = f.beta_dict(w[1])
    gap = f.beta_dict(w[2])

    # canonical state order: column-major over the wall array (j outer, i inner)
    H = walls.shape[0]
    def order(rec):
        return np.argsort((rec["j"] - 1) * H + (rec["i"] - 1), kind="stable")
    o = order(nu)
    states = np.stack([nu["i"][o], nu["j"][o]], axis=1)
    key = (states[:, 1] - 1) * H + (states[:, 0] - 1)

    def values(rec):
        oo = order(rec)
        k = (rec["j"][oo] - 1) * H + (rec["i"][oo] - 1)
        assert np.array_equal(k, key), "state sets differ between stored functions"
        return rec["v"][oo].copy()

    betas = np.array(sorted(traj.keys()))
    assert np.array_equal(betas, np.array(sorted(gap.keys())))
    T = np.stack([np.stack([values(g) for g in traj[b]]) for b in betas])
    G = np.stack([np.stack([values(g) for g in gap[b]]) for b in betas])
    # every stored state is a free cell, and every free cell is stored
    assert not walls[states[:, 0] - 1, states[:, 1] - 1].any()
    assert (walls == 0).sum() == len(states)
    return dict(walls=walls, goal=goal, s0=s0, randomness=randomness, gamma=gamma,
                step_cost=step_cost, reward=reward, betas=betas, states=states,
                nu=values(nu), traj=T, gap=G)


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/data"
    dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
    os.makedirs(dst, exist_ok=True)
    for name in sorted(os.listdir(src)):
        if not name.endswith(".jld2"):
            continue
        d = decode(os.path.join(src, name))
        out = os.path.join(dst, name[:-5] + ".npz")
        np.savez_compressed(out, **d)
        s0 = d["s0"]
        k = np.where((d["states"] == s0).all(axis=1))[0][0]
        print("%s: H,W=%s goal=%s s0=%s r=%g gamma=%g cost=%r reward=%r nstates=%d nu(s0)=%.15f sum(nu)=%.12f"
              % (name, d["walls"].shape, d["goal"], s0, d["randomness"], d["gamma"], d["step_cost"],
                 d["reward"], len(d["states"]), d["nu"][k], d["nu"].sum()))


if __name__ == "__main__":
    main()
