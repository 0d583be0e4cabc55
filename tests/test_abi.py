""" The C-ABI library loads and exports exactly what include/deepcv_b200.h declares (no compute calls: no GPU needed). """
import ctypes
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _header_symbols():
    text = (ROOT / 'include' / 'deepcv_b200.h').read_text()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return set(re.findall(r'\b(dcv_[a-z0-9_]+)\s*\(', text))


def test_header_matches_binding_table():
    from deepcv_b200._lib import SYMBOLS
    assert _header_symbols() == set(SYMBOLS), (_header_symbols() ^ set(SYMBOLS))


def test_library_exports_every_declared_symbol():
    from deepcv_b200._lib import SYMBOLS, library_path
    path = library_path()
    if not path.exists():
        import __graft_entry__
        __graft_entry__.build()
    cdll = ctypes.CDLL(str(path))
    for name in _header_symbols():
        assert hasattr(cdll, name), f'{name} declared in include/deepcv_b200.h but not exported by {path.name}'
    cdll.dcv_abi_version.restype = ctypes.c_int
    from deepcv_b200._lib import ABI_VERSION
    assert cdll.dcv_abi_version() == ABI_VERSION
    exported = subprocess.run(['nm', '-D', '--defined-only', str(path)], capture_output=True, text=True).stdout
    undeclared = {s for s in re.findall(r' T (dcv_[a-z0-9_]+)', exported)} - set(SYMBOLS)
    assert not undeclared, f'exported but not declared: {undeclared}'


def test_library_is_sm100a_only_and_has_no_torch_dependency():
    from deepcv_b200._lib import library_path
    needed = subprocess.run(['readelf', '-d', str(library_path())], capture_output=True, text=True).stdout
    assert 'torch' not in needed and 'c10' not in needed
    sass = subprocess.run(['/usr/local/cuda/bin/cuobjdump', '-lelf', str(library_path())], capture_output=True, text=True).stdout
    if sass.strip():
        assert 'sm_100a' in sass and not re.search(r'sm_(?!100a)\d+', sass), sass


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    import importlib
    import deepcv_b200._lib as L
    monkeypatch.setenv('DEEPCV_B200_LIB', str(tmp_path / 'nope.so'))
    fresh = L._Library()
    with pytest.raises(RuntimeError, match='no CPU or PyTorch fallback'):
        fresh.dcv_abi_version


def test_cpu_tensors_are_rejected_not_computed(default_hp):
    import torch
    from deepcv_b200.meta.base_module import DeepcvModule
    model = DeepcvModule((3, 32, 32), default_hp)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        model(torch.zeros(2, 3, 32, 32))
