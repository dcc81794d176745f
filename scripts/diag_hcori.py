"""Developer diagnostic: the hand-crafted-orientation application test's per-keypoint accounting (tests/test_gpu_parity.py) for the library
given by AFFNET_B200_LIB - for every keypoint whose frame differs from the oracle's: is it the shape (A A^T) or the orientation, and how
close was the oracle's bin decision?"""
import os, sys
import torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import affnet_oracle as O
from helpers import gold, gray_from_rgb, load_weights, match_keypoints
from affnet_b200.architectures import AffNetFast
from affnet_b200.SparseImgRepresenter import ScaleSpaceAffinePatchExtractor
W = load_weights()
aff = AffNetFast(PS=32); aff.load_state_dict(W["affnet"]); aff = aff.eval().cuda()
f = gold("graf_full.npz")
det = ScaleSpaceAffinePatchExtractor(mrSize=5.192, num_features=3000, border=5, num_Baum_iters=1, AffNet=aff)
img = gray_from_rgb(f["rgb"])
L1, _ = det(img.cuda(), do_ori=True)
L1 = L1.cpu()
oL, _, st = O.detect(img, W["affnet"], None, 3000, do_ori=True, debug=True)
sm = O.orientation_hist_bins(st["debug"]["ori"]["patches"])
top = sm.topk(2, dim=1).values
margin = (top[:, 0] - top[:, 1]) / top[:, 0]
ia, ib = match_keypoints(oL, L1)
A0, A1 = oL[ia][:, :, :2], L1[ib][:, :, :2]
sc = (A0[:, 0, 0] * A0[:, 1, 1] - A0[:, 0, 1] * A0[:, 1, 0]).abs().sqrt()
eA = (A0 - A1).abs().amax(dim=(1, 2)) / sc
M0, M1 = A0 @ A0.transpose(1, 2), A1 @ A1.transpose(1, 2)
eM = (M0 - M1).abs().amax(dim=(1, 2)) / (sc * sc)
print("matched %d of %d; max shape error (A A^T, relative) %.2e; keypoints with frame error > 1e-2: %d" % (len(ia), oL.shape[0], eM.max().item(), int((eA > 1e-2).sum())))
for i in (eA > 1e-2).nonzero().view(-1).tolist():
    ang0 = torch.atan2(A0[i, 1, 0], A0[i, 0, 0]).item(); ang1 = torch.atan2(A1[i, 1, 0], A1[i, 0, 0]).item()
    print(" keypoint %d (oracle idx %d): frame err %.3e, shape err %.3e, oracle top-2 margin %.2e, oracle bins %s, angle(a11,a21) oracle %.4f ours %.4f rad, scale %.2f" % (
        i, ia[i], eA[i].item(), eM[i].item(), margin[ia[i]].item(), sm[ia[i]].topk(3).indices.tolist(), ang0, ang1, sc[i].item()))
print("ten largest shape errors:", ["%.1e" % v for v in eM.topk(10).values.tolist()])
