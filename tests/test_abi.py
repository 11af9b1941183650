"""CPU: the C-ABI library builds, loads and exports every symbol include/tsp.h declares;
without a GPU the product path fails loudly (no CPU fallback)."""
import os
import re

import pytest

import tree_search_planning_b200 as tsp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "tsp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tsp_[a-z_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(tsp.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = tsp.load_library()
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert lib.tsp_abi_version() == 1


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(tsp.TspError) as ei:
        tsp.Engine(tsp.get_config("highway_env_ttc_flat"), max_roots=4)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tree_search_planning_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower(), "%s mentions the oracle" % f


def test_more_than_two_players_not_implemented():
    cfg = tsp.get_config("highway_env_ttc_flat", players=[0, 1, 2])
    with pytest.raises(NotImplementedError):
        tsp.Engine(cfg, max_roots=1)
