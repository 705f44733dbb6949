import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import test_gpu_ops as T
bad = 0
for it in range(40):
    for (d, c, cmvn) in [(512, 64, False), (256, 16, True), (512, 8, True)]:
        try:
            T.test_frontend_conv0_dw1(1, d, c, cmvn)
        except AssertionError as e:
            bad += 1
            print("FAIL", it, d, c, cmvn, str(e)[:80], flush=True)
print("done, failures:", bad)
