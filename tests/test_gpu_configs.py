"""GPU sanity of the other BASELINE.json configurations (sizes the oracle is too slow for in a test): the batched pipeline
runs, reports no capacity overflow, is deterministic and keeps its size-independent properties."""
import pytest
import torch

import affnet_oracle as O
from helpers import load_weights

pytestmark = pytest.mark.gpu
W = load_weights()


@pytest.fixture(scope="module")
def nets():
    from affnet_b200.architectures import AffNetFast, OriNetFast
    from affnet_b200.HardNet import HardNet
    a, o, h = AffNetFast(PS=32), OriNetFast(PS=32), HardNet()
    a.load_state_dict(W["affnet"]); o.load_state_dict(W["orinet"]); h.load_state_dict(W["hardnet"])
    return a.eval().cuda(), o.eval().cuda(), h.eval().cuda()


@pytest.mark.parametrize("B,H,Wd,K,border,n_oct", [(2, 1080, 1920, 4000, 5, 7), (1, 2160, 3840, 8000, 33, 5), (3, 480, 640, 500, 5, None)])
def test_other_configs(nets, B, H, Wd, K, border, n_oct):
    from affnet_b200.pipeline import DetectDescribePipeline
    aff, ori, hn = nets
    imgs = torch.cat([O.synthetic_image(H, Wd, 4321 + i) for i in range(B)]).cuda()
    pipe = DetectDescribePipeline(B, H, Wd, aff, hn, ori, num_features=K, border=border, do_ori=True)
    if n_oct is not None:
        import affnet_b200._lib as L
        assert L.lib().ag_pipeline_plan(pipe._h).contents.n_octaves == n_oct      # config 5: border=33 -> the 5-octave pyramid
    lafs, resp, desc, cnt = [t.clone() for t in pipe.run(imgs)]
    pipe.check()
    l2, r2, d2, c2 = pipe.run(imgs)
    torch.cuda.synchronize()
    assert torch.equal(cnt, c2) and torch.equal(lafs, l2) and torch.equal(desc, d2)
    for b in range(B):
        n = int(cnt[b])
        assert 0 < n <= K
        r = resp[b, :n]
        assert bool((r[:-1] >= r[1:]).all())
        assert (desc[b, :n].norm(dim=1) - 1).abs().max() < 1e-4
        c = lafs[b, :n, :, 2]
        assert bool((c[:, 0] >= 0).all() and (c[:, 0] <= Wd).all() and (c[:, 1] >= 0).all() and (c[:, 1] <= H).all())


def test_edge_cases(nets):
    """Empty / tiny / under-populated inputs: no crash, no stale rows, counts consistent with the oracle."""
    from affnet_b200.SparseImgRepresenter import ScaleSpaceAffinePatchExtractor
    from affnet_b200.pipeline import DetectDescribePipeline
    aff, ori, hn = nets
    # constant image: no maxima at any level (the reference raises on torch.cat([]); we return empty tensors)
    flat = torch.full((1, 1, 96, 128), 77.0).cuda()
    det = ScaleSpaceAffinePatchExtractor(mrSize=5.192, num_features=100, border=5, num_Baum_iters=1, AffNet=aff, OriNet=ori)
    dL, r = det(flat, do_ori=True)
    assert dL.shape == (0, 2, 3) and r.shape == (0,)
    assert hn(det.extract_patches_from_pyr(dL, PS=32)).shape == (0, 128)
    pipe = DetectDescribePipeline(2, 96, 128, aff, hn, ori, num_features=100, do_ori=True)
    imgs = torch.cat([flat, O.synthetic_image(96, 128, 3).cuda()])
    lafs, resp, desc, cnt = pipe.run(imgs)
    pipe.check()
    assert int(cnt[0]) == 0 and int(cnt[1]) > 0
    # fewer candidates than requested: everything that survives is returned, in the oracle's order
    small = O.synthetic_image(64, 80, 9)
    det = ScaleSpaceAffinePatchExtractor(mrSize=5.192, num_features=5000, border=5, num_Baum_iters=1, AffNet=aff, OriNet=ori)
    dL, r = det(small.cuda(), do_ori=False)
    oL, oresp, st = O.detect(small, W["affnet"], None, 5000, do_ori=False)
    assert abs(dL.shape[0] - oL.shape[0]) <= 2 and dL.shape[0] < 5000
    if dL.shape[0] == oL.shape[0]:
        assert (dL.cpu()[:, :, 2] - oL[:, :, 2]).abs().max() < 0.05          # same order (octave, level, raster), same places
