"""Developer tool: where do the warp roles of the second-generation engine (tcx_first_kernel, tcx_conv_kernel) spend their cycles?

    AG_XPROF=1 bash affnet_b200/csrc/build.sh
    AFFNET_B200_LIB=affnet_b200/lib/libaffnet_b200_xprof.so python scripts/role_prof_x.py [n_patches]

Prints, per net and kernel, the mean over CTAs of one warp of each role: loop cycles per patch and cycles waiting on each barrier."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from affnet_b200.architectures import AffNetFast, OriNetFast  # noqa: E402
from affnet_b200.HardNet import HardNet  # noqa: E402
from affnet_b200 import _lib  # noqa: E402
from helpers import load_weights  # noqa: E402

ROLES = [("mma issuer", ["p_full", "c1_empty", "full(L1 epi done)", "tempty"]),
         ("L2 epi (set 0)", ["tfull", "-", "-", "-"]),
         ("L1 epilogue", ["empty(L2 mma done)", "c1_full", "-", "-"]),
         ("producer", ["p_empty", "-", "-", "-"])]
CONV_ROLES = [("mma issuer", ["full (input landed)", "tempty (epilogue drained)", "-", "-"]),
              ("epilogue set 0", ["tfull (MMAs done)", "-", "-", "-"]), None,
              ("loader", ["empty (stage free)", "-", "-", "-"])]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32000
    L = _lib.lib()
    L.ag_debug_role_prof_x.restype = C.c_int
    L.ag_debug_role_prof_x.argtypes = [C.c_void_p]
    g = torch.Generator().manual_seed(1)
    P = (torch.rand(n, 1, 32, 32, generator=g) * 255).cuda()
    W = load_weights()
    nets = {"affnet": AffNetFast(PS=32), "orinet": OriNetFast(PS=32), "hardnet": HardNet()}
    for name, m in nets.items():
        m.load_state_dict(W[name])
        m = m.cuda().eval()
        for _ in range(3):
            m(P)
        torch.cuda.synchronize()
        buf = np.zeros((8, 160, 20), dtype=np.uint64)
        assert L.ag_debug_role_prof_x(buf.ctypes.data_as(C.c_void_p)) == 0     # reset
        launches = _lib.profile(lambda: m(P))
        torch.cuda.synchronize()
        assert L.ag_debug_role_prof_x(buf.ctypes.data_as(C.c_void_p)) == 0
        per_cta = n / 148.0
        print("== %s: %d patches, %.1f per CTA; launches: %s" % (name, n, per_cta, [(k, round(v, 3)) for k, v in launches]))
        for slot in range(6):
            t = buf[slot, :148].astype(np.float64)
            if t.sum() == 0:
                continue
            print(" kernel slot %d (%s)" % (slot, "tcx_first: layers 1+2" if slot == 0 else "tcx_conv layer %d" % (slot + 1)))
            for r, role in enumerate(ROLES if slot == 0 else CONV_ROLES):
                if role is None:
                    continue
                rn, wn = role
                rows = t[t[:, r * 5] > 0]
                if rows.shape[0] == 0:
                    continue
                tot = rows[:, r * 5].mean()
                line = "  %-15s total %8.0f clk/patch" % (rn, tot / per_cta)
                for i in range(4):
                    if wn[i] != "-":
                        line += " | wait %s %6.0f" % (wn[i], rows[:, r * 5 + 1 + i].mean() / per_cta)
                print(line)


if __name__ == "__main__":
    main()
