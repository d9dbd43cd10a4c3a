"""Narrow-band compaction (numpy boolean indexing order), SURVEY.md section 8a row a3.

Reference call sites: ``features[mask]`` / ``velocity[mask] = ...`` / ``u[mask] += ...`` in
lsml/core/model.py:388-394 and lsml/core/fit_job_handler.py:370-375, 504-511.
"""
import torch

from .. import _lib


def compact(mask, out=None, workspace=None):
    """Ascending row-major flat indices of the true voxels of a batched mask.

    mask: CUDA bool/uint8 tensor ``(batch, *shape)``.
    Returns ``(band_idx, item_offsets)``: ``band_idx`` int64 ``(B,)`` with
    ``index = item * voxels + local`` and ``item_offsets`` int64 ``(batch + 1,)``.
    This convenience wrapper reads the band size back (one host sync); the evolution engine
    uses the non-syncing C entry point directly.
    """
    assert mask.is_cuda and mask.dtype in (torch.bool, torch.uint8)
    mask = mask.contiguous()
    total = mask.numel()
    batch = mask.shape[0]
    voxels = total // batch if batch else 0
    dev = mask.device
    with torch.cuda.device(dev):
        idx = torch.empty(max(total, 1), dtype=torch.int64, device=dev) if out is None else out
        n_band = torch.zeros(1, dtype=torch.int64, device=dev)
        offsets = torch.zeros(batch + 1, dtype=torch.int64, device=dev)
        ws_bytes = int(_lib.lib.lsml_compact_workspace_bytes(_lib.c_i64(total)))
        ws = (torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if workspace is None
              else workspace)
        _lib.check(_lib.lib.lsml_compact_mask(
            _lib.c_i64(total), _lib.c_i64(voxels), _lib.ptr(mask), _lib.ptr(idx),
            _lib.ptr(n_band), _lib.ptr(offsets), _lib.ptr(ws), _lib.current_stream()),
            "compact_mask")
        n = int(n_band.item())
    return idx[:n], offsets
