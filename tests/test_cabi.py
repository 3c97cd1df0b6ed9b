"""CPU: libsta_b200.so loads and exports every symbol include/sta_b200.h declares (no compute without a GPU), and the
ctypes structs agree with the header's field lists."""
from __future__ import annotations

import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "sta_b200.h").read_text()


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g

    g.build()
    from diffusion_spacetime_attn_b200 import native

    return ctypes.CDLL(str(native.LIB_PATH))


def declared_functions():
    return sorted(set(re.findall(r"^(?:int|const char\*)\s+(sta_\w+)\(", HEADER, flags=re.M)))


def test_header_declares_the_documented_entry_points():
    from diffusion_spacetime_attn_b200 import native

    assert declared_functions() == sorted(native.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} is declared in include/sta_b200.h but not exported"


def test_version_and_error_string(lib):
    lib.sta_version.restype = ctypes.c_int
    lib.sta_last_error.restype = ctypes.c_char_p
    m = re.search(r"#define STA_B200_VERSION (\d+)", HEADER)
    assert lib.sta_version() == int(m.group(1))
    assert isinstance(lib.sta_last_error(), bytes)


def test_null_arguments_are_rejected_without_touching_the_gpu(lib):
    from diffusion_spacetime_attn_b200 import native

    for fn, argt in (("sta_sattn_fwd", native.SattnFwdArgs), ("sta_xattn_fwd", native.XattnFwdArgs),
                     ("sta_sattn_bwd", native.SattnBwdArgs), ("sta_xattn_bwd", native.XattnBwdArgs)):
        f = getattr(lib, fn)
        f.argtypes = [ctypes.POINTER(argt), ctypes.c_void_p]
        f.restype = ctypes.c_int
        assert f(ctypes.byref(argt()), None) == 1  # STA_ERR_BAD_ARG
        lib.sta_last_error.restype = ctypes.c_char_p
        assert b"null pointer" in lib.sta_last_error()


@pytest.mark.parametrize("struct,cname", [("SattnFwdArgs", "sta_sattn_fwd_args"), ("SattnBwdArgs", "sta_sattn_bwd_args"),
                                          ("XattnFwdArgs", "sta_xattn_fwd_args"), ("XattnBwdArgs", "sta_xattn_bwd_args"),
                                          ("ProbeArgs", "sta_probe_args"), ("GroupNormArgs", "sta_groupnorm_args"),
                                          ("AddLayerNormArgs", "sta_add_layernorm_args"),
                                          ("AddLayerNormBwdArgs", "sta_add_layernorm_bwd_args"),
                                          ("GegluArgs", "sta_geglu_args"), ("Upsample2xArgs", "sta_upsample2x_args"),
                                          ("PlmsStepArgs", "sta_plms_step_args"), ("PlmsStepBwdArgs", "sta_plms_step_bwd_args")])
def test_ctypes_structs_mirror_the_header(struct, cname):
    from diffusion_spacetime_attn_b200 import native

    body = re.search(r"typedef struct \{([^{}]*)\} " + cname + ";", HEADER, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        decl = re.sub(r"^(const\s+)?(void|float|uint8_t|int32_t|int64_t|uint32_t|uint64_t)\s*\*?\s*", "", decl)
        for part in decl.split(","):
            names.append(re.sub(r"\[.*\]", "", part).replace("*", "").strip())
    assert names == [f[0] for f in getattr(native, struct)._fields_]


STRUCTS = {"SattnFwdArgs": "sta_sattn_fwd_args", "SattnBwdArgs": "sta_sattn_bwd_args", "XattnFwdArgs": "sta_xattn_fwd_args",
           "XattnBwdArgs": "sta_xattn_bwd_args", "GroupNormArgs": "sta_groupnorm_args", "ProbeArgs": "sta_probe_args",
           "AddLayerNormArgs": "sta_add_layernorm_args", "AddLayerNormBwdArgs": "sta_add_layernorm_bwd_args",
           "GegluArgs": "sta_geglu_args", "Upsample2xArgs": "sta_upsample2x_args",
           "PlmsStepArgs": "sta_plms_step_args", "PlmsStepBwdArgs": "sta_plms_step_bwd_args"}


def test_ctypes_struct_sizes_and_offsets_match_the_c_compiler(tmp_path):
    """Field NAMES can agree while the layout drifts (a float followed by an int64 is padded): compile the header with gcc
    and compare sizeof / offsetof of every field with ctypes."""
    import shutil
    import subprocess

    from diffusion_spacetime_attn_b200 import native

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{ROOT / "include" / "sta_b200.h"}"', "int main(void) {"]
    for py, cname in STRUCTS.items():
        lines.append(f'  printf("{py} size %zu\\n", sizeof({cname}));')
        for fname, _ in getattr(native, py)._fields_:
            lines.append(f'  printf("{py} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-std=c11", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    for line in out.splitlines():
        py, field, value = line.split()
        st = getattr(native, py)
        got = ctypes.sizeof(st) if field == "size" else getattr(st, field).offset
        assert got == int(value), f"{py}.{field}: ctypes {got} != C {value}"


def test_ops_refuse_cpu_tensors():
    import torch

    from diffusion_spacetime_attn_b200 import ops

    q = torch.zeros(1, 16, 320, dtype=torch.float16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.sattn_fwd(q, q, q, heads=8)


def test_fast_div_magic_numbers_divide_exactly(tmp_path):
    """The elementwise kernels replace `index / runtime_divisor` by (umulhi(n, magic) + n) >> shift (csrc/sta_common.cuh
    fast_div) with magic / shift from csrc/sta_host.h make_fast_div_raw: compile that host function with g++ and check the
    quotient against the C division for the divisors the launchers pass (channel vectors, widths, heights) and awkward ones,
    over small, random and the largest allowed (< 2^31) dividends."""
    import shutil
    import subprocess

    gxx = shutil.which("g++")
    cuda_inc = Path("/usr/local/cuda/include")
    if gxx is None or not (cuda_inc / "cuda_runtime.h").exists():
        pytest.skip("g++ / CUDA headers not available")
    src = tmp_path / "fastdiv.cpp"
    src.write_text(f'''
#include "{ROOT / "diffusion_spacetime_attn_b200" / "csrc" / "sta_host.h"}"
#include <stdlib.h>
int main() {{
  const unsigned int divisors[] = {{1, 2, 3, 5, 7, 12, 40, 64, 80, 96, 125, 128, 160, 192, 256, 320, 640, 1000, 1280, 4095, 4096,
                                   65537, (1u << 20) + 7, (1u << 30) + 1, (1u << 31) - 1}};
  unsigned long long bad = 0, checked = 0;
  srand(1);
  for (unsigned int d : divisors) {{
    unsigned int magic, shift;
    sta::make_fast_div_raw(d, &magic, &shift);
    auto check = [&](unsigned int n) {{
      const unsigned int q = (unsigned int)((((unsigned long long)n * magic) >> 32) + n) >> shift;  // 32-bit add, as on the device
      bad += q != n / d;
      ++checked;
    }};
    for (unsigned int n = 0; n < 70000; ++n) check(n);
    for (int i = 0; i < 200000; ++i) check((((unsigned int)rand() << 16) ^ (unsigned int)rand()) & 0x7fffffffu);
    for (unsigned int n = 0x7fffffffu; n > 0x7fffffffu - 1000; --n) check(n);
    for (unsigned int k = 1; k < 2000 && (unsigned long long)k * d < (1ull << 31); ++k) {{ check(k * d); check(k * d - 1); }}
  }}
  printf("%llu %llu\\n", bad, checked);
  return bad != 0;
}}
''')
    exe = tmp_path / "fastdiv"
    subprocess.run([gxx, "-std=c++17", "-O1", f"-I{cuda_inc}", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    bad, checked = out.stdout.split()
    assert out.returncode == 0 and int(bad) == 0 and int(checked) > 1_000_000, out.stdout
