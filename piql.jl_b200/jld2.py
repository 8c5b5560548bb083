"""Writer for the reference's result files: `@save "data/<name>.jld2" out` (scripts/distributed_experiments.jl:49-59).

The reference stores one NamedTuple `out = (ν, traj, gap, randomness, γ, ctrl, s0)` per configuration with JLD2.jl, an HDF5
container: a 512-byte text header, a version-2 superblock at byte 512 (all addresses relative to it), version-2 object
headers, datasets in compact (or, beyond 64 KB, contiguous) layout, and 8-byte object references between them.  This module
writes that container and that object graph from a `run_experiment` result of the Python mirror (solvers.run_experiment):

    out (216-byte struct of 27 words: refs to ν / traj / gap, randomness, γ, the closure captures of ctrl -- action range,
         walls, step cost, reward, goal -- and s0; the layout tools/make_golden.py documents and decodes)
    ν    -> Growable -> Vector{Pair{CartesianIndex{2},Float64}} of {i::Int64, j::Int64, v::Float64} records (1-based)
    traj -> Dict{Float64,Vector{Growable}}: 8 x {β, ref -> 6 refs -> Growable}          gap: the same shape
    walls (Array{Bool,2}, column-major bytes, HDF5 dims reversed), goal (2 x Int64)

What it is: the same bytes-level structure `tools/make_golden.py` reads from the reference's own files, with HDF5 datatype
messages and Jenkins lookup3 checksums as the format specifies, so a result computed on the GPU round-trips through the
decoder that pins the oracle (tests/test_jld2_writer.py).  What it is not: JLD2.jl additionally commits Julia type
descriptors (`_types/...` with attributes naming `PIQL.Growable`, the closure types of `ctrl`, ...); those are not
reproduced, so JLD2.jl would reconstruct anonymous structs rather than `PIQL.Growable`.  Neither libhdf5 nor Julia exists in
this image, so compatibility beyond the round trip is untested.  Under Julia the shim's `run_experiment` returns real
`PIQL.Growable`s and `@save` itself writes the file.
"""
import struct

import numpy as np

BASE = 512
UNDEF = 0xFFFFFFFFFFFFFFFF
HEADER = b"HDF5-based Julia Data Format, version 0.1.1 (Julia 1.10.0 64-bit LE)"


# ---- Jenkins lookup3 (hashlittle), the checksum of HDF5 version-2 metadata ---------------------------------------------
def _rot(x, k):
    return ((x << k) | (x >> (32 - k))) & 0xFFFFFFFF


def lookup3(data, initval=0):
    a = b = c = (0xDEADBEEF + len(data) + initval) & 0xFFFFFFFF
    k, n = 0, len(data)
    while n > 12:
        a = (a + int.from_bytes(data[k:k + 4], "little")) & 0xFFFFFFFF
        b = (b + int.from_bytes(data[k + 4:k + 8], "little")) & 0xFFFFFFFF
        c = (c + int.from_bytes(data[k + 8:k + 12], "little")) & 0xFFFFFFFF
        a = (a - c) & 0xFFFFFFFF; a ^= _rot(c, 4); c = (c + b) & 0xFFFFFFFF
        b = (b - a) & 0xFFFFFFFF; b ^= _rot(a, 6); a = (a + c) & 0xFFFFFFFF
        c = (c - b) & 0xFFFFFFFF; c ^= _rot(b, 8); b = (b + a) & 0xFFFFFFFF
        a = (a - c) & 0xFFFFFFFF; a ^= _rot(c, 16); c = (c + b) & 0xFFFFFFFF
        b = (b - a) & 0xFFFFFFFF; b ^= _rot(a, 19); a = (a + c) & 0xFFFFFFFF
        c = (c - b) & 0xFFFFFFFF; c ^= _rot(b, 4); b = (b + a) & 0xFFFFFFFF
        k += 12; n -= 12
    if n == 0:
        return c
    tail = data[k:] + b"\0" * (12 - n)
    a = (a + int.from_bytes(tail[0:4], "little")) & 0xFFFFFFFF
    b = (b + int.from_bytes(tail[4:8], "little")) & 0xFFFFFFFF
    c = (c + int.from_bytes(tail[8:12], "little")) & 0xFFFFFFFF
    c ^= b; c = (c - _rot(b, 14)) & 0xFFFFFFFF
    a ^= c; a = (a - _rot(c, 11)) & 0xFFFFFFFF
    b ^= a; b = (b - _rot(a, 25)) & 0xFFFFFFFF
    c ^= b; c = (c - _rot(b, 16)) & 0xFFFFFFFF
    a ^= c; a = (a - _rot(c, 4)) & 0xFFFFFFFF
    b ^= a; b = (b - _rot(a, 14)) & 0xFFFFFFFF
    c ^= b; c = (c - _rot(b, 24)) & 0xFFFFFFFF
    return c


# ---- HDF5 datatype messages (class | version << 4, 3 bytes of class bits, size, properties) --------------------------------
def dt_int(size, signed=True):
    return struct.pack("<BBBBI", 0x10 | 0, 0x08 if signed else 0x00, 0, 0, size) + struct.pack("<HH", 0, 8 * size)


def dt_f64():
    return struct.pack("<BBBBI", 0x10 | 1, 0x20, 63, 0, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)


def dt_ref():
    return struct.pack("<BBBBI", 0x10 | 7, 0, 0, 0, 8)


def dt_compound(members, size):
    """version-3 compound: members = [(name, byte offset, datatype message)]"""
    body = b""
    width = 1 if size < 256 else 2 if size < 65536 else 4
    for name, off, dt in members:
        body += name.encode() + b"\0" + off.to_bytes(width, "little") + dt
    return struct.pack("<BBBBI", 0x30 | 6, len(members) & 0xFF, len(members) >> 8, 0, size) + body


DT_RECORD = dt_compound([("i", 0, dt_int(8)), ("j", 8, dt_int(8)), ("v", 16, dt_f64())], 24)
DT_BETA = dt_compound([("first", 0, dt_f64()), ("second", 8, dt_ref())], 16)
DT_REFBOX = dt_compound([("ref", 0, dt_ref())], 8)
DT_PAIR = dt_compound([("i", 0, dt_int(8)), ("j", 8, dt_int(8))], 16)
DT_OUT = dt_compound([("w%d" % k, 8 * k, dt_int(8, signed=False)) for k in range(27)], 216)


class Writer:
    """Appends objects after the superblock and hands out their (base-relative) addresses."""

    def __init__(self):
        self.buf = bytearray(HEADER.ljust(BASE, b"\0"))
        self.buf += b"\0" * 48                       # superblock, filled in by finish()

    def _addr(self):
        while len(self.buf) % 8:
            self.buf += b"\0"
        return len(self.buf) - BASE

    def raw(self, data):
        a = self._addr()
        self.buf += data
        return a

    def object(self, messages):
        """version-2 object header from [(type, payload)]; returns its address."""
        body = b"".join(struct.pack("<BHB", t, len(p), 0) + p for t, p in messages)
        flags = 2                                    # 4-byte chunk size, no times, no attribute phase change
        head = b"OHDR" + struct.pack("<BB", 2, flags) + struct.pack("<I", len(body)) + body
        a = self._addr()
        self.buf += head + struct.pack("<I", lookup3(head))
        return a

    def dataset(self, dims, dtype, data):
        """dims: HDF5 order (slowest first); () writes a scalar dataspace."""
        if len(dims):
            space = struct.pack("<BBBB", 2, len(dims), 0, 1) + b"".join(struct.pack("<Q", d) for d in dims)
        else:
            space = struct.pack("<BBBB", 2, 0, 0, 0)
        if len(data) <= 60000:
            layout = struct.pack("<BBH", 3, 0, len(data)) + bytes(data)                 # compact
        else:
            layout = struct.pack("<BBQQ", 3, 1, self.raw(bytes(data)), len(data))       # contiguous
        return self.object([(0x01, space), (0x03, dtype), (0x08, layout)])

    def group(self, links):
        msgs = [(0x02, struct.pack("<BBQQ", 0, 0, UNDEF, UNDEF)), (0x0A, struct.pack("<BB", 0, 0))]
        for name, addr in links.items():
            n = name.encode()
            msgs.append((0x06, struct.pack("<BBB", 1, 0, len(n)) + n + struct.pack("<Q", addr)))
        return self.object(msgs)

    def finish(self, root):
        sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBB", 2, 8, 8, 0) + struct.pack("<4Q", BASE, UNDEF, len(self.buf), root)
        sb += struct.pack("<I", lookup3(sb))
        self.buf[BASE:BASE + len(sb)] = sb
        return bytes(self.buf)


def _records(w, values, walls):
    """Growable over the free cells of `walls` ((H, W) array, NaN-free at free cells) -> address of the Growable object."""
    free = np.argwhere(np.asarray(walls).T == 0)[:, ::-1]                        # column-major order of (i, j)
    rec = np.empty(len(free), dtype=np.dtype([("i", "<i8"), ("j", "<i8"), ("v", "<f8")]))
    rec["i"], rec["j"] = free[:, 0] + 1, free[:, 1] + 1                           # Julia indices are 1-based
    rec["v"] = np.asarray(values, dtype=np.float64)[free[:, 0], free[:, 1]]
    arr = w.dataset((len(rec),), DT_RECORD, rec.tobytes())
    return w.dataset((), DT_REFBOX, struct.pack("<Q", arr))                       # Growable{...}: ref -> Dict slots


def _beta_dict(w, betas, series, walls):
    pairs = []
    for beta, vs in zip(betas, series):
        refs = [_records(w, v, walls) for v in vs]
        vec = w.dataset((len(refs),), dt_ref(), struct.pack("<%dQ" % len(refs), *refs))
        pairs.append(w.dataset((), DT_BETA, struct.pack("<dQ", float(beta), vec)))
    slots = w.dataset((len(pairs),), dt_ref(), struct.pack("<%dQ" % len(pairs), *pairs))
    return w.dataset((), DT_REFBOX, struct.pack("<Q", slots))


def _dense(g, walls):
    a = getattr(g, "array", None)
    if a is None:
        a = g.dense(walls) if hasattr(g, "dense") else np.asarray(g)
    return np.asarray(a, dtype=np.float64)


def save_result(path, out, step_cost=None, reward=None):
    """Writes `out` -- the dict solvers.run_experiment returns: nu, traj {β: [V_0..V_n]}, gap {β: [...]}, randomness, gamma,
    ctrl, s0 (0-based state) -- as data/<name>.jld2 of the reference.  step_cost / reward default to the values probed from
    ctrl.reward_function (what gridworld_ctrl captured)."""
    ctrl = out["ctrl"]
    gw = ctrl.grid
    walls = np.asarray(gw.walls) != 0
    H, W = walls.shape
    goal = tuple(int(v) for v in gw.goal)
    if step_cost is None or reward is None:
        other = next((int(i), int(j)) for i, j in np.argwhere(~walls) if (int(i), int(j)) != goal)
        reward = float(ctrl.reward_function(other, 0, goal))
        step_cost = -float(ctrl.reward_function(other, 0, other))
    betas = sorted(out["traj"].keys())
    w = Writer()
    nu = _records(w, _dense(out["nu"], walls), walls)
    traj = _beta_dict(w, betas, [[_dense(v, walls) for v in out["traj"][b]] for b in betas], walls)
    gap = _beta_dict(w, betas, [[_dense(v, walls) for v in out["gap"][b]] for b in betas], walls)
    wall_obj = w.dataset((W, H), dt_int(1, signed=False), np.asfortranarray(walls.astype(np.uint8)).T.tobytes())
    goal_obj = w.dataset((), DT_PAIR, struct.pack("<2q", goal[0] + 1, goal[1] + 1))
    rnd, gam = float(out["randomness"]), float(out["gamma"])
    f = lambda x: struct.unpack("<Q", struct.pack("<d", x))[0]
    s0 = out["s0"]
    words = [nu, traj, gap, f(rnd), f(gam), 0, 3, 4, f(rnd), wall_obj, f(step_cost), f(reward), wall_obj, goal_obj,
             f(rnd), 4, wall_obj, f(rnd), 4, wall_obj, goal_obj, goal_obj, wall_obj, goal_obj, f(gam),
             int(s0[0]) + 1, int(s0[1]) + 1]
    assert len(words) == 27
    out_obj = w.dataset((), DT_OUT, struct.pack("<27Q", *words))
    root = w.group({"out": out_obj})
    data = w.finish(root)
    with open(path, "wb") as fh:
        fh.write(data)
    return path
