"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/nerfb200.h
declares, its config struct matches the Python mirror, and — with no GPU — every constructor fails loudly
instead of falling back to a CPU path."""
import ctypes
import re
from pathlib import Path

import pytest

import nerf_or_nothing_b200 as nb

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    hdr = (ROOT / "include" / "nerfb200.h").read_text()
    names = set(re.findall(r"\b(nerf_[a-z0-9_]+)\s*\(", hdr))
    names.discard("nerf_output_gradient_cb")
    return names


def test_library_exports_every_declared_symbol():
    l = nb.lib()
    names = _declared()
    assert len(names) >= 50
    missing = [n for n in sorted(names) if not hasattr(l, n)]
    assert not missing, missing
    assert names == set(nb.EXPORTED_SYMBOLS)


def test_config_struct_layout_and_defaults():
    c = nb.default_config()
    # reference compile-time constants: ANU/helpers.h:16-20, ANU/AcceleratedMLP.h:10-19
    assert (c.n_rays, c.n_samples, c.n_levels) == (1024, 128, 2)
    assert (c.net_depth, c.net_width, c.net_depth_condition, c.net_width_condition, c.skip_layer) == (8, 256, 1, 128, 4)
    assert (c.deg_point, c.deg_view) == (16, 4)
    assert abs(c.coarse_loss_mult - 0.1) < 1e-7 and abs(c.resample_padding - 0.01) < 1e-7
    assert c.density_bias == 0.0 and c.rgb_padding == 0.0 and c.white_bkgd == 1
    assert c.seed == 7 and c.engine_flags == 0 and ctypes.sizeof(c) == 104


def test_header_cites_reference_for_every_group():
    hdr = (ROOT / "include" / "nerfb200.h").read_text()
    for cite in ("ANU/AcceleratedMipNeRF.cpp:52-144", "ANU/AcceleratedMLP.cpp:214-255", "ANU/AcceleratedMLP.cpp:256-321",
                 "ANU/AcceleratedAdamOptimizer.h:5-20", "ANU/AcceleratedGradientCalculator.h:8-17",
                 "ANU/OutputRetriever.h:7-11", ".cu:318-344", ".cu:362-402", ".cu:403-416", ".cu:187-221", ".cu:292-317"):
        assert cite in hdr, cite


def test_no_cpu_fallback_without_gpu():
    n = ctypes.c_int()
    nb.lib().nerf_device_count(ctypes.byref(n))
    if n.value > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(nb.NerfError, match="no CPU path"):
        nb.AcceleratedMipNeRF()
    with pytest.raises(nb.NerfError, match="no CPU path"):
        nb.AcceleratedAdamOptimizer([16, 4])
    with pytest.raises(nb.NerfError, match="no CPU path"):
        nb.AcceleratedGradientCalculator(8)
    import numpy as np
    with pytest.raises(nb.NerfError, match="no CPU path"):
        nb.BinDataset(np.zeros((4, 16), np.float32))        # the resident dataset lives in device memory only
    with pytest.raises(nb.NerfError):
        nb.BinDataset("/nonexistent/train_data.bin")         # a missing file is an error, not an empty dataset
    # per-stage entry points refuse too (null pointers are never dereferenced without a device)
    assert nb.lib().nerf_adam_optimizer_step(None, None, None, None, 0.1, 0.9, 0.999, 1.0, 1.0, 4, 0) == 100002


def test_product_never_imports_the_oracle():
    for p in (ROOT / "nerf_or_nothing_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".cpp", ".h") and p.is_file():
            txt = p.read_text()
            assert "oracle/" not in txt.replace("test oracle", "") or "never" in txt or "oracle draws" in txt, p
            assert "import oracle" not in txt and "from oracle" not in txt, p


def test_lr_schedule_matches_oracle():
    from oracle import oracle as orc

    for step in (0, 1, 100, 2500, 50000, 1000000):
        assert abs(nb.learning_rate_decay(step) - orc.learning_rate_decay(step)) <= 1e-6 * orc.learning_rate_decay(step) + 1e-12


def _build_cpp_example(tmp_path):
    import subprocess

    exe = tmp_path / "train_loop"
    libdir = ROOT / "nerf_or_nothing_b200"
    cmd = ["/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++", "-std=c++17", "-O1", f"-I{ROOT / 'include'}",
           str(ROOT / "examples" / "train_loop.cpp"), f"-L{libdir}", "-lnerfb200", f"-Wl,-rpath,{libdir}", "-o", str(exe)]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def test_cpp_host_mirror_compiles_links_and_fails_loudly_without_gpu(tmp_path):
    """include/nerfb200.hpp (the five reference class names in C++) links against the C ABI; with no GPU the native
    host exits with the library's 'no CPU path' error instead of computing anything."""
    import subprocess

    nb.lib()
    exe = _build_cpp_example(tmp_path)
    n = ctypes.c_int()
    nb.lib().nerf_device_count(ctypes.byref(n))
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    if n.value == 0:
        assert r.returncode == 2 and "no CPU path" in r.stderr, (r.returncode, r.stderr)
    else:
        assert r.returncode == 0 and "Step 3/3, Loss:" in r.stdout, (r.returncode, r.stdout, r.stderr)


def test_header_is_plain_c_and_every_entry_links_from_c(tmp_path):
    """The drop-in boundary is a C ABI: include/nerfb200.h must compile as C99 (what a P/Invoke, cgo or ctypes binding
    assumes: no C++ types, no overloads, no default arguments) and every declared entry must resolve when a C host links
    against libnerfb200.so.  The host only takes addresses and calls the two device-free entries."""
    import subprocess

    nb.lib()
    names = sorted(_declared())
    src = ['#include "nerfb200.h"', "#include <stdio.h>", "int main(void) {", "  const void* syms[] = {"]
    src += [f"    (const void*)&{n}," for n in names]
    src += ["  };", "  nerf_config c; nerf_default_config(&c);",
            '  printf("%d %d %d %d\\n", (int)(sizeof(syms) / sizeof(syms[0])), nerf_version(), c.n_samples, (int)sizeof(nerf_config));',
            "  return 0;", "}"]
    c_file, exe = tmp_path / "abi_c99.c", tmp_path / "abi_c99"
    c_file.write_text("\n".join(src) + "\n")
    libdir = ROOT / "nerf_or_nothing_b200"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", f"-I{ROOT / 'include'}", str(c_file), f"-L{libdir}", "-lnerfb200",
                        f"-Wl,-rpath,{libdir}", "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60).stdout.split()
    assert out == [str(len(names)), "100", "128", str(ctypes.sizeof(nb.default_config()))], out


def test_engine_flag_values_agree_between_header_and_python():
    """nerf_config.engine_flags: every NERF_FLAG_* of include/nerfb200.h has the same value as nb.FLAG_*, and the bits are distinct."""
    import re

    import nerf_or_nothing_b200 as nb

    text = (ROOT / "include" / "nerfb200.h").read_text()
    flags = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define NERF_FLAG_(\w+)\s+(\d+)u", text)}
    assert len(flags) >= 10 and len(set(flags.values())) == len(flags)
    assert all(v and v & (v - 1) == 0 for v in flags.values())  # single bits
    for name, value in flags.items():
        assert getattr(nb, "FLAG_" + name) == value, name
