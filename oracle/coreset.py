"""Test infrastructure (CPU restatement): coreset selection over ConvNet3D embeddings, following
distill_coreset.py:75-110 of the reference step by step (python loops, small inputs only)."""
import torch


def k_center(features, ipc):
    """distill_coreset.py:79-90: the sample nearest to the class mean, then greedy farthest-point additions.  The reference reduces
    the (n,) centre distances over the sample axis at :87 (`torch.min(dis_center, dim=-1)` on a vector), so its second centre is
    always sample 0, and from the third centre on :86 no longer broadcasts and raises.  This restatement follows the intended rule
    (distance to the NEAREST chosen centre, argmax over samples); it is pinned against the reference where the reference is well
    defined (first centre, tests/golden/coreset.npz)."""
    mean = features.mean(dim=0, keepdim=True)
    order = torch.argsort(torch.norm(features - mean, dim=1))
    chosen = [int(order[0])]
    for _ in range(ipc - 1):
        d = torch.stack([torch.norm(features - features[c], dim=-1) for c in chosen], 0)
        chosen.append(int(torch.argmax(d.min(dim=0).values)))
    return chosen


def herding(features, ipc):
    """distill_coreset.py:97-109: greedy choice so that the running sum tracks (i+1) * mean."""
    mean = features.mean(dim=0, keepdim=True)
    chosen, left = [], list(range(features.shape[0]))
    for i in range(ipc):
        target = mean * (i + 1) - (features[chosen].sum(dim=0) if chosen else 0)
        j = int(torch.argmin(torch.norm(target - features[left], dim=1)))
        chosen.append(left.pop(j))
    return chosen
