"""Host-side expansion micro-benchmark (no GPU needed): mapf_unpack_records on 1 Mi random packed agents for 1..16
threads and each instruction-set variant (MAPF_HOST_ISA=generic|bmi2|vbmi caps it).

usage: PYTHONPATH=. python tools/bench_host_unpack.py
Measured on the B200 host (Xeon, 16 cores, AVX-512 VBMI): vbmi 3.5 ns per agent on one thread, 0.49 ms per Mi agents
on 16 threads; bmi2 4.1 ns / 0.55 ms.  (The pool is created inside every call here; mapf_step_host keeps it.)
"""
import ctypes as C, time, numpy as np, os, sys
from dl_reference_models_b200 import _native as nat
L = nat.lib()
n, v2 = 1 << 20, 25
rs = L.mapf_packed_record_bytes(v2)
rng = np.random.default_rng(0)
packed = rng.integers(0, 256, n * rs + 64, dtype=np.uint8)
o = np.zeros((n, v2), np.uint8); m = np.zeros((n, 5), np.int8); gd = np.zeros((n, 2), np.float32); rw = np.zeros(n, np.float32)
ptr = lambda a: a.ctypes.data_as(C.c_void_p)
for th in (1, 4, 8, 16):
    best = 1e9
    for _ in range(7):
        t0 = time.perf_counter()
        L.mapf_unpack_records(ptr(packed), n, v2, th, ptr(o), ptr(m), ptr(gd), ptr(rw), 31.0, 31.0)
        best = min(best, time.perf_counter() - t0)
    print(os.environ.get('MAPF_HOST_ISA', 'auto'), th, 'threads: %.3f ms, %.2f ns/agent/thread' % (best * 1e3, best * 1e9 * th / n))
