"""Coreset baselines over the fast embed path (reference: distill_coreset.py:75-110): k-center and herding selection of
``ipc`` real videos per class from ConvNet3D embeddings.  The embeddings come from the tensor-core path (or the exact fp32
kernels); the selection itself is a few tiny device reductions per class."""
import torch


def k_center_select(features, ipc):
    """Indices of the sample nearest to the class mean followed by greedy farthest-point additions (:79-90).  From the second
    centre on this is the INTENDED rule: the reference reduces over the wrong axis (:87) and always appends sample 0, then raises
    for ipc > 2 (see oracle/coreset.py)."""
    mean = features.mean(dim=0, keepdim=True)
    first = torch.argsort(torch.norm(features - mean, dim=1))[0]
    chosen = [int(first)]
    dmin = torch.norm(features - features[first], dim=-1)               # distance to the nearest chosen centre
    for _ in range(ipc - 1):
        nxt = torch.argmax(dmin)
        chosen.append(int(nxt))
        dmin = torch.minimum(dmin, torch.norm(features - features[nxt], dim=-1))
    return chosen


def herding_select(features, ipc):
    """Greedy mean matching (:97-109): pick the remaining sample closest to (i+1)*mean - sum(chosen)."""
    mean = features.mean(dim=0, keepdim=True)
    n = features.shape[0]
    left = torch.ones(n, dtype=torch.bool, device=features.device)
    running = torch.zeros_like(mean)
    chosen = []
    for i in range(ipc):
        dis = torch.norm(mean * (i + 1) - running - features, dim=1)
        dis = torch.where(left, dis, torch.full_like(dis, float('inf')))
        j = int(torch.argmin(dis))                                        # first minimum = the reference's order in idx_left
        chosen.append(j)
        left[j] = False
        running = running + features[j]
    return chosen


def select_coreset(embed_fn, videos, labels, num_classes, ipc, method='k-center', chunk=256):
    """``embed_fn(videos_chunk) -> (n, D)``; returns (image_syn (C*ipc, ...), label_syn, chosen dataset indices)."""
    labels = torch.as_tensor(labels).long().cpu()
    pick = {'k-center': k_center_select, 'herding': herding_select}
    if method not in pick:
        raise NotImplementedError(method)
    out, idx_all = [], []
    for c in range(num_classes):
        idx = torch.nonzero(labels == c).flatten()
        imgs = videos[idx.to(videos.device)]
        with torch.no_grad():
            feats = torch.cat([embed_fn(imgs[s:s + chunk]) for s in range(0, imgs.shape[0], chunk)], 0).float()
        sel = pick[method](feats, ipc)
        out.append(imgs[sel])
        idx_all += [int(idx[j]) for j in sel]
    image_syn = torch.cat(out, 0)
    label_syn = torch.arange(num_classes, device=image_syn.device).repeat_interleave(ipc)
    return image_syn, label_syn, idx_all
