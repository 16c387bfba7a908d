"""CPU: host logic of the flow trainers that needs no device -- the eager -> capture -> replay sequencing of run() and
the per-block strided-view arithmetic FlowTrainStep uses to address one tensor of all coupling blocks at once."""
import torch

from links_b200.flowtrain import _NAMES, _GraphReplay, _rup


class _FakeGraph:
    def __init__(self, log):
        self.log = log

    def replay(self):
        self.log.append("replay")


class _Fake(_GraphReplay):
    def __init__(self, ok=True):
        self.log, self.ok = [], ok

    def step(self):
        self.log.append("eager")

    def capture(self, warmup=1):
        self.log.append("capture(warmup=%d)" % warmup)
        self.graph = _FakeGraph(self.log)
        return self.graph

    def _graph_ok(self):
        return self.ok


def test_run_sequences_eager_capture_replay():
    t = _Fake()
    for _ in range(4):
        t.run()
    assert t.log == ["eager", "capture(warmup=0)", "replay", "replay", "replay"]
    t = _Fake()
    for _ in range(3):
        t.run(use_graph=False)                 # --no-graph
    assert t.log == ["eager"] * 3 and t.graph is None
    t = _Fake(ok=False)                        # data parallel: the NCCL all-reduce stays an eager call
    for _ in range(3):
        t.run()
    assert t.log == ["eager"] * 3 and t.graph is None


def test_all_blocks_strided_views_address_the_per_block_slots():
    """The flat parameter buffer holds, block after block, the six trainable tensors of a coupling block, each padded to a
    multiple of 64 floats: view [k] of the as_strided tensor FlowTrainStep builds must be block k's own view."""
    C_dim, nb, hidden = 34, 8, 1024
    c1, c2 = C_dim - C_dim // 2, C_dim // 2
    shapes = {"subnet.0.weight": (hidden, c1), "subnet.0.bias": (hidden,), "subnet.2.weight": (2 * c2, hidden),
              "subnet.2.bias": (2 * c2,), "global_scale": (1, C_dim), "global_offset": (1, C_dim)}
    numel = {n: int(torch.tensor(shapes[n]).prod()) for n in _NAMES}
    span = sum(_rup(numel[n], 64) for n in _NAMES)
    flat = torch.arange(span * nb, dtype=torch.float32)
    first, o = {}, 0
    for n in _NAMES:
        first[n] = o
        o += _rup(numel[n], 64)
    per_block, off = [], 0
    for k in range(nb):
        d = {}
        for n in _NAMES:
            d[n] = flat[off:off + numel[n]].view(shapes[n])
            off += _rup(numel[n], 64)
        per_block.append(d)
    assert off == span * nb
    for n in _NAMES:
        allb = torch.as_strided(flat, (nb,) + tuple(shapes[n]), (span,) + tuple(torch.empty(shapes[n]).stride()), first[n])
        for k in range(nb):
            assert allb[k].data_ptr() == per_block[k][n].data_ptr()
            assert torch.equal(allb[k], per_block[k][n])
