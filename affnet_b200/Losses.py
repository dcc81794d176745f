"""Losses.py counterpart on the evaluation side of the hot path: distance_matrix_vector (Losses.py:5-13) and the
second-nearest-neighbour ratio matcher of train_AffNet_test_on_graffity.py:292-298."""
import torch

from . import _lib as L


def distance_matrix_vector(anchor, positive):
    """[n1,D], [n2,D] CUDA float32 -> [n1,n2] sqrt(|a|^2 + |b|^2 - 2 a.b + 1e-6)."""
    a, b = L.f32c(anchor, "anchor"), L.f32c(positive, "positive")
    out = torch.empty(a.size(0), b.size(0), dtype=torch.float32, device=a.device)
    if a.size(0) and b.size(0):
        L.check(L.lib().ag_distance_matrix(L.ptr(a), a.size(0), L.ptr(b), b.size(0), a.size(1), L.ptr(out), L.stream_ptr()))
    return out


def match_snn(descriptors1, descriptors2, SNN_threshold=0.8):
    """-> (tent_matches_in_1, tent_matches_in_2, min_dist, min_2nd_dist) exactly as the reference's test() computes them,
    including `dist_matrix[:, idxs_in_2] = 100000` before the second minimum."""
    a, b = L.f32c(descriptors1, "descriptors1"), L.f32c(descriptors2, "descriptors2")
    n1, n2 = a.size(0), b.size(0)
    dev = a.device
    if a.dim() != 2 or b.dim() != 2 or (n1 and n2 and a.size(1) != b.size(1)):
        raise L.AffnetB200Error("match_snn: descriptors must be [n1,D] and [n2,D] with the same D")
    if n1 == 0 or n2 == 0:     # the reference returns empty matches
        e = torch.empty(0, dtype=torch.long, device=dev)
        return e, e.clone(), torch.empty(n1, device=dev), torch.empty(n1, device=dev)
    nb = L.lib().ag_match_snn_workspace_bytes(n1, n2)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    idx = torch.empty(n1, dtype=torch.int32, device=dev)
    mn = torch.empty(n1, dtype=torch.float32, device=dev)
    sec = torch.empty(n1, dtype=torch.float32, device=dev)
    keep = torch.empty(n1, dtype=torch.uint8, device=dev)
    L.check(L.lib().ag_match_snn(L.ptr(a), n1, L.ptr(b), n2, a.size(1), float(SNN_threshold), L.ptr(ws), nb, L.ptr(idx), L.ptr(mn), L.ptr(sec),
                                 L.ptr(keep), L.stream_ptr()))
    mask = keep.bool()
    return torch.arange(n1, device=dev)[mask], idx.long()[mask], mn, sec
