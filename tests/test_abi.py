"""CPU tier: the C-ABI library builds for sm_100a, loads without a GPU, and exports exactly the
symbols include/oake_b200.h declares (no compute calls here)."""
import ctypes
import pathlib
import re
import subprocess
import threading

import pytest

from oadp_b200 import binding, build

ROOT = pathlib.Path(__file__).resolve().parent.parent
HEADER = ROOT / 'include' / 'oake_b200.h'


def declared_symbols():
    text = re.sub(r'/\*.*?\*/', '', HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r'\b(oake_[a-z0-9_]+)\s*\(', text)))


def test_header_matches_binding():
    assert declared_symbols() == sorted(binding.SIGNATURES)


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.oake_abi_version() == 1
    assert lib.oake_act_dtype() in (b'f16', b'bf16')
    # the error string is per thread: a thread that has not made a failing call sees none
    seen = []
    t = threading.Thread(target=lambda: seen.append(lib.oake_last_error()))
    t.start()
    t.join()
    assert seen == [b'']


def test_sass_is_blackwell_native():
    so = build.build()
    sass = subprocess.run(['cuobjdump', '-sass', str(so)], capture_output=True, text=True).stdout
    assert 'sm_100a' in sass
    for mnemonic in ('UTCHMMA', 'UTMALDG', 'LDTM'):  # tcgen05.mma, TMA load, tcgen05.ld
        assert mnemonic in sass, mnemonic


def test_create_rejects_bad_arguments_without_gpu(lib):
    w = binding.Weights()
    h = ctypes.c_void_p()
    assert lib.oake_create(ctypes.byref(h), 0, ctypes.byref(w)) != 0
    assert b'layers' in lib.oake_last_error()
    assert lib.oake_workspace_bytes(None, 1, 0, None) != 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(binding, '_lib', None)
    monkeypatch.setattr(binding, 'LIB_PATH', tmp_path / 'nope.so')
    with pytest.raises(binding.OakeError):
        binding.load()


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under oadp_b200/ or the oadp/ alias package may import it
    (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs do)."""
    offenders = []
    for pkg in ('oadp_b200', 'oadp'):
        for f in (ROOT / pkg).rglob('*.py'):
            if re.search(r'^\s*(from|import)\s+oracle\b', f.read_text(), flags=re.M):
                offenders.append(str(f.relative_to(ROOT)))
    assert offenders == []
    bench = (ROOT / 'bench.py').read_text()
    # in bench.py the oracle is only reachable from the two BASELINE functions (CPU oracle, library GPU
    # baseline), never from the timed product path in main()
    imports = [m.start() for m in re.finditer(r'^\s*(?:from|import)\s+oracle\b', bench, flags=re.M)]
    assert len(imports) == 3
    cpu_fn, lib_fn, main_fn = bench.index('def cpu_reference('), bench.index('def library_baseline('), bench.index('def main(')
    assert cpu_fn < lib_fn < main_fn
    assert all(cpu_fn < i < main_fn for i in imports)
    assert sum(1 for i in imports if i > lib_fn) == 1
