"""CPU-side checks of the C ABI: the library loads, exports every symbol include/ood_b200.h declares, the ctypes
struct mirrors match the C layout, and argument errors are reported without touching a GPU."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'ood_b200.h')


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as g
    g.build()
    from ood_gan_inversion_b200 import _lib
    return _lib


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(ood_[a-z0-9_]+)\s*\(', src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = declared_functions()
    assert len(names) >= 20
    handle = C.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f'{n} declared in ood_b200.h but not exported'
        assert n in lib.EXPORTS, f'{n} has no ctypes signature in _lib.py'
    assert lib.lib().ood_version() == 100


def test_struct_layout_matches_c(lib, tmp_path):
    prog = tmp_path / 'layout.c'
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ood_b200.h"\nint main(void){\n'
                    'printf("%zu %zu %zu %zu\\n", sizeof(ood_conv3x3_args), offsetof(ood_conv3x3_args, noise_bstride),'
                    ' offsetof(ood_conv3x3_args, batch), offsetof(ood_conv3x3_args, out_f32));\n'
                    'printf("%zu %zu %zu %zu\\n", sizeof(ood_blur_act_args), offsetof(ood_blur_act_args, out_img),'
                    ' offsetof(ood_blur_act_args, taps), offsetof(ood_blur_act_args, dtype));\nreturn 0;}\n')
    exe = tmp_path / 'layout'
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), str(prog), '-o', str(exe)])
    out = subprocess.check_output([str(exe)], text=True).split()
    a, b = lib.ConvArgs, lib.BlurActArgs
    assert [int(v) for v in out[:4]] == [C.sizeof(a), a.noise_bstride.offset, a.batch.offset, a.out_f32.offset]
    assert [int(v) for v in out[4:]] == [C.sizeof(b), b.out_img.offset, b.taps.offset, b.dtype.offset]


def test_argument_errors_are_reported_without_a_gpu(lib):
    L = lib.lib()
    rc = L.ood_upfirdn2d(None, None, None, 1, 4, 4, 4, 4, 1, 1, 1, 1, 0, 0, 0, 0, 0, None)
    assert rc == -1 and b'null' in L.ood_last_error()
    rc = L.ood_fused_bias_act(None, None, None, None, 16, 4, 4, 1, 0, 0.2, 1.0, 0, None)
    assert rc == -1 and b'act=3' in L.ood_last_error()
    a = lib.ConvArgs()
    assert L.ood_conv3x3(C.byref(a), None) == -1
    with pytest.raises(RuntimeError):
        lib.check(-1, 'demo')


def test_missing_library_fails_loudly(lib, monkeypatch):
    monkeypatch.setattr(lib, '_lib', None)
    monkeypatch.setattr(lib, 'LIB_PATH', '/nonexistent/libood_b200.so')
    with pytest.raises(RuntimeError, match='no CPU or PyTorch fallback'):
        lib.lib()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'ood_gan_inversion_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py') and fn != 'smoke.py':          # smoke() is a named checker leg
            assert 'oracle' not in open(os.path.join(pkg, fn)).read(), fn
