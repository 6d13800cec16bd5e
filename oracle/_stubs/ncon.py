"""Stand-in for the un-vendored third-party `ncon` package (reference dependency:
setup.cfg:33, requirements.txt:5; version unpinned; absent from /root/reference and from
this image, no network).  TEST INFRASTRUCTURE ONLY: it exists so that the unmodified
reference (`/root/reference/src/tnsu`) can be imported by `oracle/make_golden.py`.

Restates ncon's published algorithm (Pfeifer et al., arXiv:1402.0939): positive labels are
contracted, negative labels stay open.  Walk the contraction `order`; take the first pending
label, find the tensors that carry it, contract ALL labels those two tensors share in one
`tensordot` (or a partial trace when one tensor carries the label twice), append the result, and
continue; finally combine leftovers by outer product and permute the open legs to `forder`.
The reference's single call site is simple_update.py:584-589; it is pinned by the reference's
own tests/test_main.py:17-35 (value reproduced with 0.0 error, tests/golden/manifest.json).
"""
import numpy as np


def ncon(tensors, indices, order=None, forder=None):
    tensors = [np.asarray(t) for t in tensors]
    labels = [[int(x) for x in idx] for idx in indices]
    if order is None:
        order = sorted({x for idx in labels for x in idx if x > 0})
    order = [int(x) for x in order]
    if forder is None:
        forder = sorted({x for idx in labels for x in idx if x < 0}, reverse=True)
    forder = [int(x) for x in forder]

    while order:
        lab = order[0]
        holders = [n for n, idx in enumerate(labels) if lab in idx]
        if len(holders) == 1:  # partial trace inside one tensor
            n = holders[0]
            idx = labels[n]
            dup = [x for x in set(idx) if idx.count(x) == 2 and x in order]
            t = tensors[n]
            for x in dup:
                a, b = [p for p, y in enumerate(idx) if y == x]
                t = np.trace(t, axis1=a, axis2=b)
                idx = [y for y in idx if y != x]
            tensors[n], labels[n] = t, idx
            done = dup
        else:
            a, b = holders[0], holders[1]
            shared = [x for x in labels[a] if x in labels[b]]
            ax_a = [labels[a].index(x) for x in shared]
            ax_b = [labels[b].index(x) for x in shared]
            t = np.tensordot(tensors[a], tensors[b], axes=(ax_a, ax_b))
            idx = [x for x in labels[a] if x not in shared] + [x for x in labels[b] if x not in shared]
            for n in sorted((a, b), reverse=True):
                del tensors[n], labels[n]
            tensors.append(t)
            labels.append(idx)
            done = shared
        order = [x for x in order if x not in done]

    out, idx = tensors[0], labels[0]
    for t, i in zip(tensors[1:], labels[1:]):  # disconnected pieces: outer product
        out = np.multiply.outer(out, t)
        idx = idx + i
    return np.transpose(out, [idx.index(x) for x in forder])
