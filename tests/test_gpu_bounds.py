"""Out-of-bounds canaries (compute-sanitizer is closed on this pool): every array handed to the C ABI
is carved out of one guarded allocation with 0xAB-filled gaps before and after it; after running
every kernel on ragged sizes the gaps must be untouched."""
import ctypes as C

import pytest
import torch

from merging_gym_b200 import _native as nat

pytestmark = pytest.mark.gpu
GUARD = 512


class Arena:
    def __init__(self, nbytes):
        self.buf = torch.full((nbytes,), 0xAB, dtype=torch.uint8, device="cuda")
        self.off = GUARD
        self.spans = []

    def take(self, nbytes, fill=None):
        start = (self.off + 255) // 256 * 256
        self.spans.append((start, nbytes))
        self.off = start + nbytes + GUARD
        assert self.off + GUARD <= self.buf.numel()
        v = self.buf[start:start + nbytes]
        if fill is not None:
            v.fill_(fill)
        return v

    def check(self):
        mask = torch.ones(self.buf.numel(), dtype=torch.bool, device="cuda")
        for s, n in self.spans:
            mask[s:s + n] = False
        assert bool((self.buf[mask] == 0xAB).all()), "a kernel wrote outside its arrays"


@pytest.mark.parametrize("n", [1, 63, 77, 130, 1000])
@pytest.mark.parametrize("random_reset", [False, True])
def test_no_kernel_writes_out_of_bounds(n, random_reset):
    lib = nat.load()
    K = 5
    ar = Arena(64 << 20)
    f64 = [ar.take(8 * n, 0) for _ in range(6)]
    meta = ar.take(4 * n, 0)
    obs = ar.take(40 * n * K); rew = ar.take(8 * n * K); done = ar.take(n * K); info = ar.take(n * K)
    term = ar.take(40 * n); epr = ar.take(8 * n); epl = ar.take(4 * n)
    a1 = ar.take(n, 2); a2 = ar.take(n, 3); acts = ar.take(2 * n * K)
    a64 = ar.take(8 * n, 0)
    stats = ar.take(8 * nat.STATS_ROWS * nat.STATS_COLS, 0)
    ring = ar.take(4 * 22 * 50, 0); counter = ar.take(8, 0); ids = ar.take(4 * 50, 0)
    scratch = ar.take(4 * int(lib.mg_record_scratch_words(n)), 0)
    act_out = ar.take(n); q_out = ar.take(4 * 5 * n)
    w1t = ar.take(4 * 10 * 200, 0); b1 = ar.take(4 * 200, 0); w2t = ar.take(4 * 200 * 4 * 28, 0)
    b2 = ar.take(4 * 100, 0); w3 = ar.take(4 * 5 * 100, 0); b3 = ar.take(4 * 5, 0)
    p = lambda t: C.c_void_p(t.data_ptr())
    st = nat.MgState(*[t.data_ptr() for t in f64], meta.data_ptr())
    out = nat.MgOut(obs.data_ptr(), rew.data_ptr(), done.data_ptr(), info.data_ptr(),
                    term.data_ptr(), epr.data_ptr(), epl.data_ptr())
    rs = nat.MgResetSpec(nat.RESET_RANDOM if random_reset else nat.RESET_FIXED, 0, 7, 1 << 40)
    rw = nat.default_rewards()
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    nat.check(lib.mg_reset(C.byref(st), n, None, p(obs), 0, C.byref(rs), s), "reset")
    for t in range(60):
        nat.check(lib.mg_sample_actions(p(a1), p(a2), n, 5, 0, t, s), "sample")
        nat.check(lib.mg_step(C.byref(st), n, p(a1), p(a2), nat.ACT_U8, C.byref(rw), C.byref(out), p(stats), 1,
                              C.byref(rs), s), "step")
    nat.check(lib.mg_step(C.byref(st), n, p(a1), None, nat.ACT_U8, C.byref(rw), C.byref(out), p(stats), 0,
                          C.byref(rs), s), "step pve")
    nat.check(lib.mg_step(C.byref(st), n, p(a64), p(a64), nat.ACT_I64, C.byref(rw), C.byref(out), None, 1,
                          C.byref(rs), s), "step i64")
    for pvp in (1, 0):
        nat.check(lib.mg_rollout(C.byref(st), n, pvp, 5, 0, 100, K, C.byref(rw), C.byref(out), p(acts), p(stats), 1,
                                 C.byref(rs), s), "rollout")
    nat.check(lib.mg_reset(C.byref(st), n, p(a1), p(obs), 0, C.byref(rs), s), "masked reset")
    nat.check(lib.mg_record_transitions(p(obs), p(obs), p(term), p(a1), p(a2), p(rew), p(done), p(info), None, None, n, 1, 0, 1,
                                        p(ring), 50, p(ids), p(counter), p(scratch), s), "record")
    nat.check(lib.mg_record_transitions(p(obs), p(obs), None, p(a1), None, p(rew), p(done), p(info), None, None, n, 0, 1, 2,
                                        p(ring), 50, None, p(counter), p(scratch), s), "record log")
    nat.check(lib.mg_record_transitions(p(obs), p(obs), p(term), p(a1), p(a2), p(rew), p(done), p(info), p(a1), p(a2), n, 0, 2, 1,
                                        p(ring), 45, None, p(counter), p(scratch), s), "record hdqn rows")
    for flags in (0, 1):                                        # plain and mirrored (opponent-view) read
        nat.check(lib.mg_mlp_act(p(obs), None, n, 10, 5, p(w1t), p(b1), p(w2t), p(b2), p(w3), p(b3), p(act_out),
                                 p(q_out), flags, s), "mlp")
    w2tc = ar.take(4 * 2 * 25 * 14 * 2 * 8 * 4, 0)
    for flags in (0, 1):
        nat.check(lib.mg_mlp_act_tc(p(obs), None, n, 10, 5, p(w1t), p(b1), p(w2tc), p(b2), p(w3), p(b3), p(act_out),
                                    p(q_out), flags, s), "mlp tc")
    blob = ar.take(64 + 13312 + 93184, 0)                        # the packed fp16 operands of MG_MLP_FLAG_F16X3 (all zero)
    for flags in (0, 1):
        nat.check(lib.mg_mlp_act_tc(p(obs), None, n, 10, 5, None, None, p(blob), p(b2), p(w3), p(b3), p(act_out),
                                    p(q_out), flags | nat.MLP_FLAG_F16X3, s), "mlp tc f16x3")
    torch.cuda.synchronize()
    ar.check()
    assert int(counter.view(torch.int64)[0]) > 0
