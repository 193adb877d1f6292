import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
osb = importlib.import_module("optimization-solvers_b200")
n = 16384
names = {0: "gemv (read n^2)", 1: "eager update (R+W n^2)", 2: "gemvT", 3: "cudaMemcpyAsync D2D (R+W n^2)"}
names.update({10:'gridstride U4 g=8x', 11:'gridstride U8 g=8x', 12:'slab U8 g=1x', 13:'slab U8 g=2x', 14:'slab U4 g=4x', 15:'copy slab U8 g=2x (out of place)', 16:'gridstride U8 g=16x', 17:'slab U16 g=1x'})
names.update({18:'gridstride U8 512thr/SM', 19:'gridstride U8 1024thr/SM', 20:'gridstride U16 1024thr/SM', 21:'gridstride U16 2048thr/SM', 22:'gridstride U4 g=16x'})
for which in (3, 18, 19, 11, 20, 21, 22, 16):
    ms = osb.bench_qn_kernel(which, n, reps=30)
    byt = n * n * 8 * (1 if which in (0, 2) else 2)
    print("%-34s %.4f ms  %.1f GB/s" % (names[which], ms, byt / ms / 1e6))
