"""CPU: the C-ABI library builds/loads and exports every symbol include/probpose_b200.h declares;
the product path refuses to run without a GPU instead of falling back to the CPU."""

import ctypes
import re
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    from probpose_pytorch_b200.build import build
    path = build()
    return ctypes.CDLL(str(path))


def _declared():
    text = (ROOT / "include" / "probpose_b200.h").read_text()
    return sorted(set(re.findall(r"^PP_API\s+[\w\s\*]+?\b(pp_\w+)\(", text, flags=re.M)))


def test_header_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/probpose_b200.h but not exported"


def test_binding_covers_header():
    from probpose_pytorch_b200 import _lib
    assert sorted(_lib.EXPORTS) == _declared()


def test_version_and_error_string(lib):
    lib.pp_version.restype = ctypes.c_int
    lib.pp_last_error_string.restype = ctypes.c_char_p
    assert lib.pp_version() == 2
    lib.pp_source_hash.restype = ctypes.c_char_p
    from probpose_pytorch_b200.build import source_hash
    assert lib.pp_source_hash().decode() == source_hash()   # the binary was compiled from the sources in the tree
    assert isinstance(lib.pp_last_error_string(), bytes)


def test_invalid_arguments_return_status_not_crash(lib):
    # argument validation happens before any CUDA call, so this is safe without a GPU
    lib.pp_last_error_string.restype = ctypes.c_char_p
    assert lib.pp_encode(None, None, None, None, None, None, None, None, None) == -1
    assert b"pp_encode" in lib.pp_last_error_string()
    assert lib.pp_decode_expected(None, None, None, None, None, None, None, None, None, ctypes.c_int64(0), None) == -1
    assert lib.pp_heatmap_tail(None, None, 0, ctypes.c_int64(4), ctypes.c_float(0.5), None) == -1


def test_struct_layouts_match_header():
    from probpose_pytorch_b200 import _lib
    assert ctypes.sizeof(_lib.EncodeParams) == 44
    assert ctypes.sizeof(_lib.DecodeParams) == 48
    assert ctypes.sizeof(_lib.LossParams) == 72
    assert ctypes.sizeof(_lib.OksTable) == 56


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    import probpose_pytorch_b200 as pp
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pp.get_heatmap_maximum(np.zeros((2, 4, 4), dtype=np.float32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pp.ProbMap((192, 256), (48, 64), np.full(17, 0.05)).encode(np.zeros((1, 17, 2)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pp.OKSHeatmapLoss()(torch.zeros(1, 1, 4, 4), torch.zeros(1, 1, 4, 4))


def test_product_does_not_import_oracle():
    pkg = ROOT / "probpose_pytorch_b200"
    for f in pkg.rglob("*.py"):
        assert not re.search(r"^\s*(from|import)\s+oracle\b", f.read_text(), flags=re.M), f


def test_host_tables_match_oracle():
    """Host constant tables (product) against the oracle's restatement of the reference."""
    import oracle as oc
    from probpose_pytorch_b200 import _tables, synth
    for wl in (synth.WORKLOADS[1], synth.WORKLOADS[4], synth.WORKLOADS[5]):
        W, H = wl.heatmap_size
        K = wl.num_keypoints
        assert np.array_equal(_tables.oks_variance(wl.sigmas, H, W), oc.oks_variance_table(wl.sigmas, H, W))
        assert np.array_equal(_tables.encode_divisors(wl.sigmas, -1, K, H, W), 2 * oc.oks_variance_table(wl.sigmas, H, W))
        assert np.array_equal(_tables.encode_divisors(wl.sigmas, 2.0, K, H, W), np.full(K, 4.0))
        ref = oc.oks_kernels_2d(K, H, W, wl.sigmas)
        s = _tables.oks_variance(wl.sigmas, H, W)
        for k in (0, K // 2, K - 1):
            r = int(np.ceil(3 * s[k]))
            ax = np.arange(-r, r + 1)
            dist = np.sqrt(ax[None, :] ** 2 + ax[:, None] ** 2)
            ker = np.exp(-(dist ** 2) / (2 * s[k]))
            assert np.array_equal(ker / ker.sum(), ref[k])
    assert np.array_equal(_tables.gaussian_taps(11), oc.gaussian_taps_f32(11))


def test_header_is_plain_c_and_links_from_c(lib, tmp_path):
    """The boundary is a C ABI: the header compiles as C99 and a C program links against the library."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    root = Path(__file__).resolve().parents[1]
    so = root / "probpose_pytorch_b200" / "csrc" / "libprobpose_b200.so"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-x", "c",
                    str(root / "include" / "probpose_b200.h")], check=True)
    exe = tmp_path / "cabi_smoke"
    subprocess.run([gcc, "-std=c99", "-Wall", "-I", str(root / "include"), str(root / "tests" / "c" / "cabi_smoke.c"),
                    "-o", str(exe), str(so), f"-Wl,-rpath,{so.parent}"], check=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "sizeof encode/decode/loss params: 44 48 72" in res.stdout
    from probpose_pytorch_b200 import _lib
    import ctypes
    assert f"sizeof mailbox: {ctypes.sizeof(_lib.Mailbox)}" in res.stdout      # the ctypes mirror has the C layout


def test_loss_out_and_unit_upstream_argument_checks():
    """Host-side argument handling of the small-step helpers (no GPU needed): `loss_out` must be a one-element float32
    tensor on the prediction's device; `unit_upstream` hands out one cached scalar 1 per (device, dtype)."""
    import pytest
    import torch
    from probpose_pytorch_b200 import unit_upstream
    from probpose_pytorch_b200.loss import _check_loss_out, _is_unit_upstream
    out = torch.zeros(2, 3, 4, 4)
    assert _check_loss_out(None, out) is None
    assert _check_loss_out(torch.zeros(1), out).shape == (1,)
    assert _check_loss_out(torch.zeros(()), out).shape == (1,)
    for bad in (torch.zeros(2), torch.zeros(1, dtype=torch.float64)):
        with pytest.raises(ValueError):
            _check_loss_out(bad, out)
    one = unit_upstream("cpu")
    assert one is unit_upstream(torch.device("cpu")) and float(one) == 1.0 and one.dim() == 0
    assert _is_unit_upstream(one) and _is_unit_upstream(one.reshape(1))
    assert not _is_unit_upstream(torch.ones(())) and not _is_unit_upstream(one.clone())
