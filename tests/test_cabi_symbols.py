"""CPU test: the C-ABI library builds, loads, and exports every symbol include/b2t_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported(pkg):
    hdr = open(os.path.join(ROOT, "include", "b2t_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(b2t_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 20
    lib = ctypes.CDLL(os.path.join(ROOT, "nejm-brain-to-text_b200", "libb2t_b200.so"))
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing


def test_layout_matches_reference_state_dict(pkg):
    import b2t_pkg
    E = b2t_pkg.submodule("engine")
    cfg = E.make_config(512, 768, 5, 45, 41, 14, 4, 0.4, 0.2)
    lay = E.param_layout(cfg)
    assert sum(r * c for _, _, r, c in lay) == 44315177            # SURVEY.md section 6 (model size)
    names = [n for n, _, _, _ in lay]
    assert names[0] == "day_weights.0" and "gru.weight_ih_l0" in names and "gru.bias_hh_l4" in names and names[-1] == "h0"
    shapes = {n: (r, c) for n, _, r, c in lay}
    assert shapes["gru.weight_ih_l0"] == (2304, 7168) and shapes["gru.weight_hh_l3"] == (2304, 768) and shapes["out.weight"] == (41, 768)
    assert all(off % 64 == 0 for _, off, _, _ in lay)


def test_no_cpu_path(pkg):
    """The product must fail loudly without a GPU rather than fall back."""
    import pytest
    import torch
    import b2t_pkg
    E = b2t_pkg.submodule("engine")
    cfg = E.make_config(32, 64, 1, 2, 41, 14, 4)
    flat = torch.zeros(E.param_elems(cfg))
    with pytest.raises(Exception):
        E.Engine(cfg, flat, max_batch=2, max_T=30, training=False)
