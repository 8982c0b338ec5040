"""Loads the ctypes stub printed in INTEGRATION.md (section 2: the file a maintainer of the reference would add) as a
module, with the library name replaced by the in-tree path -- so that the documented binding is executed by the
tests instead of rotting.  Test infrastructure."""
import os
import re
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def stub_source():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    src = next(b for b in blocks if "def make(" in b)
    assert 'C.CDLL("libskyjo_b200.so")' in src
    return src.replace('"libskyjo_b200.so"', repr(os.path.join(ROOT, "skyjo_rl_b200", "libskyjo_b200.so")))


def load_stub():
    m = types.ModuleType("rlskyjo_skyjo_b200_stub")
    exec(compile(stub_source(), "INTEGRATION.md#stub", "exec"), m.__dict__)
    return m
