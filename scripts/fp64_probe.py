import ctypes as C, numpy as np, sys
sys.path.insert(0, '.')
from fetching_b200 import _lib
lib = _lib.load()
out = np.zeros(16)
_lib.check(lib.pf_debug_fp64_probe(0, out.ctypes.data_as(C.POINTER(C.c_double))), "probe")
print("TFMA/s (x2 = TFLOP/s); rows: warps/SM 4,8,16,32; cols: ILP 1,2,4,8")
print(np.round(out.reshape(4, 4), 3))
clk = 1.965e9; sms = 148
print("FMA/clk/SM:"); print(np.round(out.reshape(4,4) * 1e12 / clk / sms, 2))

out4 = np.zeros(4)
_lib.check(lib.pf_debug_issue_probe(0, out4.ctypes.data_as(C.POINTER(C.c_double))), "issue probe")
print("DFMA/clk/SM with k integer IMADs interleaved per DFMA (k=0..3):", np.round(out4 * 1e12 / clk / sms, 2))

out3 = np.zeros(5)
_lib.check(lib.pf_debug_rf_probe(0, out3.ctypes.data_as(C.POINTER(C.c_double))), "rf probe")
print("DFMA/clk/SM with 1, 2, 3 distinct register operands:", np.round(out3[:3] * 1e12 / clk / sms, 2))
print("2-register DFMA + 2 / 4 one-register ALU instructions each:", np.round(out3[3:] * 1e12 / clk / sms, 2))

out8 = np.zeros(8)
_lib.check(lib.pf_debug_dmma_probe(0, out8.ctypes.data_as(C.POINTER(C.c_double))), "dmma probe")
print("DMMA m8n8k4 FMA/clk/SM; rows: warps/SM 8,16; cols: accumulator tiles per warp 1,2,4,8:")
print(np.round(out8.reshape(2, 4) * 1e12 / clk / sms, 2))

out8 = np.zeros(8)
_lib.check(lib.pf_debug_dfma_dmma_probe(0, out8.ctypes.data_as(C.POINTER(C.c_double))), "dfma+dmma probe")
print("clocks per block per sub-partition warp, (DFMA, DMMA) = (16,0) (0,2) (0,4) (16,1) (16,2) (16,4) (32,4) (8,4):")
print(np.round(out8, 1))
