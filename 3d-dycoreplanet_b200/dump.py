"""Reader / writer of the cross-validation container declared in include/dcp_dump.h.

`DumpProblem` exposes a dump through the interface of `harness.Problem` (`P[name]`, `P.scalar(name)`, `P.names()`,
`P.spec`, `P.dim`, `P.n_cells`, `P.csr(name)`), so everything that runs on the stand-in harness -- the oracle, the
device model, the parity tests -- runs unchanged on arrays a real deal.II build of the reference produced."""
import struct

import numpy as np

MAGIC = b"DCPDUMP1"
_DTYPES = {0: np.float64, 1: np.int32, 2: np.int64, 3: np.int8, 4: np.int16}
_CODES = {np.dtype(np.float64): 0, np.dtype(np.int32): 1, np.dtype(np.int64): 2, np.dtype(np.int8): 3, np.dtype(np.int16): 4}


def write_dump(path, arrays, scalars=None, spec=None):
    """arrays: {name: numpy array of f64 / i32 / i64 / i8}; scalars: {name: int}; spec: dict of key -> value."""
    with open(path, "wb") as f:
        f.write(MAGIC)

        def put(name, a):
            a = np.ascontiguousarray(a)
            code = _CODES[a.dtype]
            nb = name.encode()
            head = struct.pack("<i", len(nb)) + nb + struct.pack("<iq", code, a.size)
            payload = a.tobytes()
            f.write(head + payload + b"\0" * ((8 - (len(head) + len(payload)) % 8) % 8))

        for name, a in arrays.items():
            put(name, a)
        for name, v in (scalars or {}).items():
            put("scalar:" + name, np.array([v], dtype=np.int64))
        if spec is not None:
            text = ",".join(f"{k}={v}" for k, v in spec.items())
            put("spec", np.frombuffer(text.encode(), dtype=np.int8))


def read_dump(path):
    """-> (arrays, scalars, spec)"""
    arrays, scalars, spec = {}, {}, {}
    with open(path, "rb") as f:
        data = f.read()
    if data[:8] != MAGIC:
        raise ValueError(f"{path}: not a DCPDUMP1 file")
    o = 8
    while o < len(data):
        (ln,) = struct.unpack_from("<i", data, o)
        name = data[o + 4:o + 4 + ln].decode()
        code, count = struct.unpack_from("<iq", data, o + 4 + ln)
        head = 4 + ln + 12
        dt = np.dtype(_DTYPES[code])
        nbytes = count * dt.itemsize
        a = np.frombuffer(data, dtype=dt, count=count, offset=o + head).copy()
        o += head + nbytes + (8 - (head + nbytes) % 8) % 8
        if name.startswith("scalar:"):
            scalars[name[7:]] = int(a[0])
        elif name == "spec":
            text = a.tobytes().decode()
            spec = dict(kv.split("=", 1) for kv in text.split(",") if "=" in kv)
        else:
            arrays[name] = a
    return arrays, scalars, spec


def dump_problem(P, path, extra=None):
    """Write every array and scalar of a harness Problem (plus `extra` arrays, e.g. reference results)."""
    arrays = {n: P[n] for n in P.names()}
    arrays.update(extra or {})
    scalars = {n: P.scalar(n) for n in P.scalar_names()}
    write_dump(path, arrays, scalars, dict(P.spec))


class DumpProblem:
    """A dump seen through the harness.Problem interface."""

    def __init__(self, path):
        self._arrays, self._scalars, spec = read_dump(path)
        self.spec = {k: (int(v) if v.lstrip("-").isdigit() else v) for k, v in spec.items()}
        self.dim = int(self._scalars.get("dim", self.spec.get("dim", 3)))
        self.n_cells = int(self._scalars["n_cells"])

    def names(self):
        return list(self._arrays)

    def scalar_names(self):
        return list(self._scalars)

    def __getitem__(self, name):
        return self._arrays[name]

    def __contains__(self, name):
        return name in self._arrays

    def scalar(self, name):
        return self._scalars[name]

    def csr(self, name):
        rp, col = self[name + ".rowptr"], self[name + ".col"]
        return rp, col, self.scalar(name + ".n_rows"), self.scalar(name + ".n_cols")
