"""Host-side evaluation helpers of the registration workflow ("next" rows N2 / N3 of SURVEY.md section 8(f)):
`dice` (reference voxelmorph/py/utils.py:265-287) and `jacobian_determinant` (py/utils.py:473-516), same arguments and
results.  They run on numpy arrays after the registration step, off the GPU hot path — exactly where the reference runs
them (scripts/tf/test.py:76-121, fold counting after scripts/torch/register.py) — and are checked against the oracle
restatements and the live reference in tests/test_utils.py.  Device versions belong to the N3 kernel work."""
import numpy as np

__all__ = ["dice", "jacobian_determinant", "count_folds"]


def dice(array1, array2, labels=None, include_zero=False):
    """Dice overlap per label between two integer label maps; float64 vector in ascending label order.

    One pass over the voxels (a joint histogram) instead of three comparisons per label."""
    a1, a2 = np.asarray(array1).ravel(), np.asarray(array2).ravel()
    if a1.shape != a2.shape:
        raise ValueError("dice: label maps differ in size (%d vs %d voxels)" % (a1.size, a2.size))
    present = np.union1d(np.unique(a1), np.unique(a2))
    labels = present if labels is None else np.asarray(labels)
    if not include_zero:
        labels = labels[labels != 0]
    # ranks of every voxel's label among the labels that occur (exact for any integer or float label values)
    r1, r2 = np.searchsorted(present, a1), np.searchsorted(present, a2)
    n = len(present)
    size1, size2 = np.bincount(r1, minlength=n), np.bincount(r2, minlength=n)
    both = np.bincount(r1[r1 == r2], minlength=n)
    out = np.zeros(len(labels), dtype=np.float64)
    pos = np.searchsorted(present, labels)
    for i, (lab, p) in enumerate(zip(labels, pos)):
        if p < n and present[p] == lab:
            out[i] = 2.0 * both[p] / max(float(size1[p] + size2[p]), np.finfo(float).eps)
    return out


def _gradient_unit(x, axis):
    """d/d(axis) with unit spacing: central differences inside, one-sided at the two ends (np.gradient's default)."""
    g = np.empty_like(x)
    sl = [slice(None)] * x.ndim

    def at(s):
        sl[axis] = s
        return tuple(sl)

    g[at(slice(1, -1))] = (x[at(slice(2, None))] - x[at(slice(None, -2))]) * 0.5
    g[at(0)] = x[at(1)] - x[at(0)]
    g[at(-1)] = x[at(-1)] - x[at(-2)]
    return g


def jacobian_determinant(disp):
    """det J of x -> x + disp(x) for a displacement field of shape (*vol_shape, nb_dims), nb_dims in (2, 3).

    The derivative of the identity grid is the identity, so only the displacement is differentiated."""
    disp = np.asarray(disp, dtype=np.float64)
    nd = disp.ndim - 1
    if nd not in (2, 3) or disp.shape[-1] != nd:
        raise AssertionError("flow has to be 2D or 3D")
    if min(disp.shape[:-1]) < 2:
        raise ValueError("jacobian_determinant: every axis needs at least 2 samples")
    # J[a][..., c] = d(x_c + disp_c) / d x_a
    J = []
    for a in range(nd):
        g = _gradient_unit(disp, a)
        g[..., a] += 1.0
        J.append(g)
    if nd == 2:
        return J[0][..., 0] * J[1][..., 1] - J[1][..., 0] * J[0][..., 1]
    dx, dy, dz = J
    return (dx[..., 0] * (dy[..., 1] * dz[..., 2] - dy[..., 2] * dz[..., 1])
            - dx[..., 1] * (dy[..., 0] * dz[..., 2] - dy[..., 2] * dz[..., 0])
            + dx[..., 2] * (dy[..., 0] * dz[..., 1] - dy[..., 1] * dz[..., 0]))


def jacobian_determinant_device(flow, return_folds=False):
    """det J on the GPU for a CUDA tensor in the module layout (B, nb_dims, *vol) — the warp `VxmDense(..., registration=True)`
    returns — with np.gradient's differencing and the reference's determinant expansion (py/utils.py:473-516).  Returns a
    (B, *vol) tensor, plus the number of folded voxels (det <= 0, python int) when `return_folds`."""
    import torch
    from . import _lib
    from .layers import _dims
    _lib.require_cuda(flow, what="jacobian_determinant_device")
    flow = _lib.contig(flow.detach())
    B, C, D, H, W, nd = _dims(flow)
    if C != nd:
        raise _lib.VxmError("jacobian_determinant_device: expected %d displacement channels, got %d" % (nd, C))
    det = torch.empty((B,) + tuple(flow.shape[2:]), dtype=torch.float32, device=flow.device)
    folds = torch.zeros(1, dtype=torch.int64, device=flow.device) if return_folds else None
    _lib.check(_lib.load().vxm_jacdet(_lib.ptr(flow), _lib.ptr(det), _lib.ptr(folds), B, D, H, W, nd, _lib.stream_ptr()), "vxm_jacdet")
    return (det, int(folds.item())) if return_folds else det


def count_folds(flow):
    """Number of voxels where the deformation folds (det J <= 0).  `flow`: module-layout field (nb_dims, *vol) or
    (1, nb_dims, *vol) as `VxmDense(..., registration=True)` returns it.  CUDA tensors are evaluated on the device
    (csrc/eval_ops.cu), numpy arrays / CPU tensors on the host."""
    if hasattr(flow, "is_cuda") and flow.is_cuda:
        f = flow if flow.dim() in (4, 5) and flow.shape[1] == flow.dim() - 2 else flow.unsqueeze(0)
        return jacobian_determinant_device(f, return_folds=True)[1]
    f = np.asarray(flow.detach().cpu() if hasattr(flow, "detach") else flow, dtype=np.float64)
    if f.ndim in (4, 5) and f.shape[0] == 1 and f.shape[1] == f.ndim - 2:
        f = f[0]
    return int(np.count_nonzero(jacobian_determinant(np.moveaxis(f, 0, -1)) <= 0))
