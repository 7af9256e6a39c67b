"""GPU probe: tcgen05 descriptor start shifted by whole rows inside a swizzled TMA tile (sliding-window conv design)."""
import ctypes as C, os, subprocess, sys, torch
sys.path.insert(0, '.')
# the probe kernel is NOT part of libood_b200.so: it is built here into its own library, next to the product one (for its error helpers)
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(ROOT, 'gpurun_out', 'libood_probe.so')
os.makedirs(os.path.dirname(SO), exist_ok=True)
subprocess.check_call(['/usr/local/cuda/bin/nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-Xcompiler', '-fPIC', '-shared',
                       '-I', os.path.join(ROOT, 'ood_gan_inversion_b200', 'csrc'), os.path.join(HERE, 'probes', 'debug_umma.cu'),
                       os.path.join(ROOT, 'ood_gan_inversion_b200', 'csrc', 'core.cu'), '-o', SO])
lib = C.CDLL(SO)
fn = lib.ood_debug_umma_shift
fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
fn.restype = C.c_int
for bk in (64, 32):
    a = torch.randn(256, bk, device='cuda').bfloat16()
    b = torch.eye(bk, device='cuda').bfloat16()
    for ubo in (0, 1):
        res = []
        for shift in (0, 1, 2, 3, 5, 8, 9, 17):
            out = torch.zeros(128, bk, device='cuda')
            rc = fn(a.data_ptr(), b.data_ptr(), out.data_ptr(), bk, shift, ubo, None)
            if rc: print(C.cast(lib.ood_last_error, C.CFUNCTYPE(C.c_char_p))()); 
            torch.cuda.synchronize()
            ok = torch.equal(out, a[shift:shift + 128].float())
            res.append((shift, rc, ok))
        print(f'bk={bk} base_offset={ubo}:', res)
