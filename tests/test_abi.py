"""CPU tests of the drop-in boundary: the library builds, loads, and exports every symbol that
include/wgan_b200.h declares; without a GPU the product path fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "wgan_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wgan_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    from arshamg_scrnaseq_wgan_b200 import _lib
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/wgan_b200.h but not exported"
    assert set(syms) == set(_lib.SIGNATURES), set(syms) ^ set(_lib.SIGNATURES)


def test_struct_layout_matches_header():
    from arshamg_scrnaseq_wgan_b200 import _lib
    assert C.sizeof(_lib.WganConfig) == 64
    assert C.sizeof(_lib.WganTensorInfo) == 64 + 16 + 8
    assert _lib.WganConfig.learn_rate.offset == 32


def test_padded_ld(lib):
    assert lib.wgan_padded_ld(6605) == 6608 and lib.wgan_padded_ld(200) == 200
    assert lib.wgan_padded_ld(1) == 1 and lib.wgan_tensor_count() == 12


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from arshamg_scrnaseq_wgan_b200 import _lib, WGANConfig, WGANTrainer
    cfg = _lib.WganConfig(100, 600, 6605, 200, 32, 0, 0, 0, 1e-5, 0.9, 0.999, 1e-8, 0.2, 10.0, 0, 0)
    h = C.c_void_p()
    rc = lib.wgan_create(C.byref(cfg), C.byref(h))
    assert rc != 0 and not h.value
    assert b"no CUDA device" in lib.wgan_last_error(None) or b"CUDA" in lib.wgan_last_error(None)
    with pytest.raises(RuntimeError):
        WGANTrainer(WGANConfig(), max_batch=32)


def test_product_does_not_import_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may import / link / execute oracle/."""
    pkg = os.path.join(ROOT, "arshamg_scrnaseq_wgan_b200")
    bad = re.compile(r"^\s*(from\s+\.*oracle|import\s+oracle|#\s*include.*oracle)|importlib.*oracle|oracle/_ref", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert not bad.search(open(os.path.join(dirpath, f)).read()), f


def test_config_defaults_follow_reference():
    from arshamg_scrnaseq_wgan_b200 import WGANConfig, Adam_Prediction_Optimizer
    c = WGANConfig()
    assert (c.n_train_steps, c.batch_size, c.noise_input_size, c.inflate_to_size, c.gex_size,
            c.disc_internal_size, c.num_cells_train, c.num_cells_generate) == (10000, 32, 100, 600, 6605, 200, 1500, 500)
    assert c.learn_rate == 1e-5 and c.beta1 == 0.9 and c.beta2 == 0.999 and c.epsilon == 1e-8
    lit = WGANConfig.reference_literal()
    assert lit.lambda_gp == 0.0 and lit.n_critic == 1 and lit.step_order == "gen_first"
    o = Adam_Prediction_Optimizer()
    assert (o.learning_rate, o.beta1, o.beta2, o.epsilon, o.prediction, o.use_locking, o.name) == \
           (0.001, 0.9, 0.999, 1e-8, False, False, "Adam")


def test_integration_stub_matches_the_header():
    """INTEGRATION.md's ctypes stub declares struct wgan_config field by field and calls the step entry points with
    positional arguments: keep both in step with include/wgan_b200.h (they went stale once)."""
    import re
    from arshamg_scrnaseq_wgan_b200 import _lib
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = text[text.index("class Cfg(C.Structure)"):text.index("# module constants of the script")]
    names = re.findall(r'"(\w+)"', block)
    assert names == [n for n, _ in _lib.WganConfig._fields_], names
    stub = text[text.index("```python"):text.index("```", text.index("```python") + 10)]
    for fn in ("wgan_generator_grads", "wgan_critic_grads", "wgan_sample", "wgan_get_scalars"):
        i = stub.index("lib." + fn + "(")
        depth, j = 0, i + len("lib." + fn)
        args, cur = [], ""
        while True:                                    # split the call's top-level arguments
            ch = stub[j]
            if ch == "(":
                depth += 1
                if depth > 1:
                    cur += ch
            elif ch == ")":
                depth -= 1
                if depth == 0:
                    args.append(cur)
                    break
                cur += ch
            elif ch == "," and depth == 1:
                args.append(cur); cur = ""
            else:
                cur += ch
            j += 1
        assert len(args) == len(_lib.SIGNATURES[fn][1]), (fn, args)
