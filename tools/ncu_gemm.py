"""A handful of stand-alone launches of the shared-MLP GEMM kernels at the training step's shapes, for `ncu --set full`
(short command: ncu replays every launch ~40 times).  argv[1] = comma list of cases (default: all)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import check_gemm as cg

CASES = {
    "fwd": lambda: cg.run_tn(0, 524288, 128, 128, False, 1),
    "fwd_xf": lambda: cg.run_tn(0, 524288, 128, 128, True, 1),
    "dgrad": lambda: cg.run_tn(0, 524288, 128, 256, False, 0),
    "dgrad_epi2": lambda: cg.run_tn(0, 524288, 128, 256, False, 2),
    "wgrad": lambda: cg.run_wg(0, 524288, 256, 128, False),
    "wgrad_xf": lambda: cg.run_wg(0, 524288, 256, 128, True),
    "sa1_fwd_l1": lambda: cg.run_tn(0, 1048576, 64, 64, True, 1),
    "sa1_fwd_l2": lambda: cg.run_tn(0, 1048576, 128, 64, True, 1),
    "sa1_dgrad_l2": lambda: cg.run_tn(0, 1048576, 64, 128, False, 2),
    "sa1_dgrad_l1": lambda: cg.run_tn(0, 1048576, 64, 64, False, 2),
    "sa1_wgrad_l2": lambda: cg.run_wg(0, 1048576, 128, 64, True),
    "sa1_wgrad_l1": lambda: cg.run_wg(0, 1048576, 64, 64, True),
    "sa2_fwd_l0": lambda: cg.run_tn(0, 524288, 128, 192, False, 1),
    "sa2_fwd_l2": lambda: cg.run_tn(0, 524288, 256, 128, True, 1),
    "sa2_dgrad_l0": lambda: cg.run_tn(0, 524288, 192, 128, False, 0),
    "sa2_wgrad_l0": lambda: cg.run_wg(0, 524288, 128, 192, False),
    "fwd_x3": lambda: cg.run_tn(2, 524288, 128, 128, True, 1),
    "wgrad_x3": lambda: cg.run_wg(2, 524288, 256, 128, True),
}
for name in (sys.argv[1].split(",") if len(sys.argv) > 1 else CASES):
    print(name, CASES[name]())
