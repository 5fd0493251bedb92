"""cgb_peer_round: the engine's protocol round over peer memory (push into the receiver's double-buffered slot, flag, consume,
acknowledge).  Two "ranks" are played by two contexts with their own streams on ONE GPU, so the pointers that would be CUDA-IPC
mappings are plain device addresses here; the flags, the device-resident round counters, the slot alternation and the
back-pressure through acknowledgements are exactly what runs between GPUs (tools/epoch_bench.py under torchrun covers that)."""
import numpy as np
import pytest

from tests.util import rand_u64, to_dev, to_np

pytestmark = pytest.mark.gpu


class Rank:
    def __init__(self, torch, me, world, slot_words):
        import cognn_b200 as cg

        self.ctx = cg.Context(0, own_stream=True)
        self.me, self.world, self.slot = me, world, slot_words
        dev = torch.device("cuda", 0)
        self.flags = torch.zeros(2 * world, dtype=torch.int32, device=dev)      # [src] data flags, [world + dst] acks
        self.state = torch.zeros(4 * world + 1, dtype=torch.int32, device=dev)  # send seq, recv seq, send done, recv done, err
        self.arena = torch.zeros(world, 2 * slot_words, dtype=torch.int64, device=dev)  # [src] two slots

    def link(self, mode, peer):
        w, f, s = self.world, self.flags.data_ptr(), self.state.data_ptr()
        pf = peer.flags.data_ptr()
        if mode == "push":
            return (s + 4 * peer.me, s + 4 * (2 * w + peer.me), f + 4 * (w + peer.me), pf + 4 * self.me, 0)
        return (s + 4 * (w + peer.me), s + 4 * (3 * w + peer.me), f + 4 * peer.me, pf + 4 * (w + self.me), 1)

    def err(self):
        return int(self.state[-1].item())


@pytest.mark.parametrize("ctas", [1, 8])
def test_peer_rounds_two_ranks_one_gpu(ctas):
    import torch

    slot = 6000
    A, B = Rank(torch, 0, 2, slot), Rank(torch, 1, 2, slot)
    rng = np.random.default_rng(3)
    sizes = [(1, 7), (4096, 3), (5001, 900), (2, 2), (77, 5000), (1000, 1000), (3, 1)]  # words of the two messages per direction
    for rnd, (n0, n1) in enumerate(sizes):
        sent = {}
        recv = {}
        for X, Y in ((A, B), (B, A)):
            m0, m1 = rand_u64(rng, n0), rand_u64(rng, n1)
            sent[X.me] = (m0, m1, to_dev(m0), to_dev(m1))
            recv[Y.me] = (X.ctx.empty(n0), X.ctx.empty(n1))
        off1 = (n0 + 1) & ~1
        torch.cuda.synchronize()  # the uploads ran on torch's stream, the rounds run on the contexts' own streams
        for X, Y in ((A, B), (B, A)):  # every rank: ONE launch, its pushes first, then what it consumes
            _, _, d0, d1 = sent[X.me]
            r0, r1 = recv[X.me]
            out_base = Y.arena[X.me].data_ptr()  # my two slots inside the peer's arena
            in_base = X.arena[Y.me].data_ptr()
            X.ctx.peer_round([(d0.data_ptr(), out_base, n0, 0), (d1.data_ptr(), out_base + 8 * off1, n1, 0),
                              (in_base, r0.data_ptr(), n0, 1), (in_base + 8 * off1, r1.data_ptr(), n1, 1)],
                             [X.link("push", Y), X.link("recv", Y)], slot, ctas, X.state.data_ptr() + 4 * (4 * 2))
        A.ctx.sync()
        B.ctx.sync()
        assert A.err() == 0 and B.err() == 0
        for me, other in ((0, 1), (1, 0)):
            r0, r1 = recv[me]
            assert np.array_equal(to_np(r0), sent[other][0]) and np.array_equal(to_np(r1), sent[other][1]), rnd
        # the device-resident round counters advanced, the flags carry the round number, the done counters are back at zero
        for X in (A, B):
            st, fl = X.state.cpu().numpy(), X.flags.cpu().numpy()
            peer = 1 - X.me
            assert st[peer] == rnd + 1 and st[2 + peer] == rnd + 1 and st[4 + peer] == 0 and st[6 + peer] == 0
            assert fl[peer] == rnd + 1 and fl[2 + peer] == rnd + 1
    A.ctx.close()
    B.ctx.close()


def test_peer_round_sender_runs_two_rounds_ahead_but_not_three():
    """Back-pressure: a sender may fill both slots before the receiver consumes anything; the third push waits for the
    acknowledgement of the first."""
    import torch

    slot = 64
    A, B = Rank(torch, 0, 2, slot), Rank(torch, 1, 2, slot)
    msgs = [to_dev(np.full(16, 100 + i, dtype=np.uint64)) for i in range(3)]
    base = B.arena[0].data_ptr()
    err = A.state.data_ptr() + 4 * 8
    torch.cuda.synchronize()
    for i in range(3):
        A.ctx.peer_round([(msgs[i].data_ptr(), base, 16, 0)], [A.link("push", B)], slot, 1, err)
    import time

    time.sleep(0.2)
    # rounds 1 and 2 are in the slots, round 3 is blocked (the flag still says 2)
    assert int(B.flags[0].item()) == 2
    out = [B.ctx.empty(16) for _ in range(3)]
    for i in range(3):
        B.ctx.peer_round([(B.arena[0].data_ptr(), out[i].data_ptr(), 16, 0)], [B.link("recv", A)], slot, 1,
                         B.state.data_ptr() + 4 * 8)
    A.ctx.sync()
    B.ctx.sync()
    assert A.err() == 0 and B.err() == 0
    for i in range(3):
        assert np.array_equal(to_np(out[i]), np.full(16, 100 + i, dtype=np.uint64))
    A.ctx.close()
    B.ctx.close()
