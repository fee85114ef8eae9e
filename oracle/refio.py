"""Reader for ref_harness output containers + helpers to run the harness.
TEST INFRASTRUCTURE ONLY (see oracle/oracle.h)."""
from __future__ import annotations

import os
import struct
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
HARNESS = os.path.join(_HERE, "_ref", "ref_harness")


def have_ref() -> bool:
    return os.path.exists(HARNESS)


def read_out(path: str) -> dict:
    with open(path, "rb") as f:
        buf = f.read()
    assert buf[:8] == b"B200OUT1"
    o, out = 8, {}
    while o < len(buf):
        (nl,) = struct.unpack_from("<i", buf, o); o += 4
        name = buf[o:o + nl].decode(); o += nl
        t = chr(buf[o]); o += 1
        (cnt,) = struct.unpack_from("<q", buf, o); o += 8
        dt = np.float64 if t == "d" else np.int64
        out[name] = np.frombuffer(buf, dtype=dt, count=cnt, offset=o).copy(); o += cnt * 8
    return out


def run(cmd: str, prob, *args) -> dict:
    """Run `ref_harness <cmd> <problem> <out> args...` and return the arrays."""
    with tempfile.TemporaryDirectory() as td:
        pin, pout = os.path.join(td, "p.bin"), os.path.join(td, "o.bin")
        prob.save(pin)
        subprocess.check_call([HARNESS, cmd, pin] + ([args[0], pout] if cmd == "order" else [pout] + [str(a) for a in args]))
        return read_out(pout)


def time_lm(prob, steps: int, warmup: int, ceres: bool = False) -> dict:
    import json
    with tempfile.TemporaryDirectory() as td:
        pin = os.path.join(td, "p.bin")
        prob.save(pin)
        s = subprocess.check_output([HARNESS, "time", pin, str(steps), str(warmup), str(int(ceres))])
        return json.loads(s.decode().strip().splitlines()[-1])
