"""Input pipeline either side of the hot path: dataset arrays -> this rank's shard -> shuffled
batches staged in pinned host memory -> copied to the device on a side stream one batch AHEAD of the step that consumes
them (the reference feeds a DataLoader(num_workers=0) of float64 numpy rows through a synchronous pageable copy,
train_leg_torso_lifter.py:385-386)."""
import torch

from .shard import shard_bounds


class ArrayLoader:
    """Epochs of shuffled, equally sized batches of this rank's contiguous shard of (poses_2d [n,34], poses_3d [n,51]).

    Difference from the reference's DataLoader (train_leg_torso_lifter.py:385, drop_last=False): the final PARTIAL batch
    of an epoch is dropped -- the step objects own static buffers of one batch size and replay a captured CUDA graph.
    The permutation is redrawn every epoch, so the dropped poses differ from epoch to epoch."""

    def __init__(self, x2d, gt, batch_global, rank=0, world=1, seed=0, shuffle=True, pin=True):
        n = x2d.shape[0]
        # equal shards: every rank runs the same number of steps per epoch, so the per-step gradient all-reduces of
        # all ranks pair up (an extra step on one rank would meet another rank's validation collective and hang)
        b, e = shard_bounds(n, rank, world, equal=True)
        self.x = torch.as_tensor(x2d[b:e], dtype=torch.float32)
        self.gt = torch.as_tensor(gt[b:e], dtype=torch.float32) if gt is not None else None
        self.batch = batch_global // world
        if self.batch % 2 or self.batch < 2:
            raise ValueError("per-rank batch must be even (row pairs / split_data_left_right_3d)")
        self.shuffle, self.pin = shuffle, pin and torch.cuda.is_available()
        self.gen = torch.Generator().manual_seed(seed * 1000 + rank)

    def __len__(self):
        return self.x.shape[0] // self.batch

    def __iter__(self):
        n = self.x.shape[0]
        perm = torch.randperm(n, generator=self.gen) if self.shuffle else torch.arange(n)
        for i in range(len(self)):
            xb = self.x[perm[i * self.batch:(i + 1) * self.batch]]
            # a fresh pinned tensor per batch: the caching host allocator only recycles a block once the asynchronous
            # copies that read it have completed, so the host may run arbitrarily far ahead of the GPU
            yield xb.pin_memory() if self.pin else xb


def loader_from_dataset(ds, batch_global, rank=0, world=1, seed=0, shuffle=True):
    """ArrayLoader over a drop-in dataset object (utils.h36m_dataset_class.*): uses its whole arrays, not per-item
    __getitem__ calls."""
    return ArrayLoader(ds.data["poses_2d"], ds.data["poses_3d"], batch_global, rank, world, seed, shuffle)


class DevicePrefetcher:
    """Wraps an iterable of (pinned) host batches: batch i+1 is copied host->device on a side stream while the caller
    works on batch i.  Yields device tensors; two device buffers alternate, so a yielded tensor stays valid until the
    next-but-one iteration."""

    def __init__(self, loader, device="cuda"):
        self.loader, self.device = loader, torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self._dev = None

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        it = iter(self.loader)
        events = [torch.cuda.Event(), torch.cuda.Event()]
        main = torch.cuda.current_stream(self.device)

        def start(i):
            try:
                host = next(it)
            except StopIteration:
                return False
            if self._dev is None:
                self._dev = [torch.empty(host.shape, dtype=host.dtype, device=self.device) for _ in range(2)]
            self.stream.wait_stream(main)          # everything enqueued so far that reads buffer i&1 (batch i-2) is ordered first
            with torch.cuda.stream(self.stream):
                self._dev[i & 1].copy_(host, non_blocking=True)
                events[i & 1].record(self.stream)
            return True

        i = 0
        more = start(0)
        while more:
            more_next = start(i + 1)
            main.wait_event(events[i & 1])
            yield self._dev[i & 1]
            i += 1
            more = more_next


def normalize_head_device(raw, fixed_scale=None, root_joint=0, out=None):
    """normalize_head / normalize_head_test (reference utils/helpers.py:198-207, 222-230) on the device, fused with the
    dataset classes' transpose-flatten (utils/h36m_dataset_class.py:25-27).

    raw: CUDA fp32 tensor, [n, 17, 2] key-points (as stored in the reference's pickles) or [n, 34] = (17 x, 17 y) rows.
    fixed_scale=None divides by the mean root-to-head distance over all n poses (normalize_head); a number divides by that
    number (normalize_head_test's 145.40964, ...).  Returns [n, 34] fp32 (root-centred, scaled, * 1/10)."""
    from . import _cabi
    if not raw.is_cuda:
        raise _cabi.LinksError("normalize_head_device runs on a B200 only (utils.helpers.normalize_head is the host version)")
    raw = raw.contiguous().float()
    transposed = 0 if raw.dim() == 3 else 1
    n = raw.shape[0]
    assert raw.numel() == n * 34
    out = torch.empty(n, 34, dtype=torch.float32, device=raw.device) if out is None else out
    acc = torch.zeros(1, dtype=torch.float64, device=raw.device)
    _cabi.check(_cabi.lib().links_normalize_head(raw.data_ptr(), n, root_joint, transposed,
                                                 float(fixed_scale) if fixed_scale else 0.0, out.data_ptr(), acc.data_ptr(),
                                                 torch.cuda.current_stream().cuda_stream), "links_normalize_head")
    return out
