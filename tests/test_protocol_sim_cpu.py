"""Randomised-interleaving models of the mbarrier / named-barrier / bulk-group protocols of the two kernels whose
synchronisation is new and has not run on hardware yet (tests/pipeline_sim.py is the tiny scheduler):
 * xattn_tc_kernel (hallo_b200/csrc/xattn_tc.cu): TMA warp, MMA warp, 4 softmax/epilogue warps, S/P/O double-buffered;
 * gemm_tc_kernel<TEPI> epilogue (hallo_b200/csrc/gemm_tc.cu): MMA warp with two accumulators, two groups of four
   epilogue warps, three panel buffers per group, TMA stores with lazy read-waits, residual prefetch two panels ahead.
Each agent is a generator that mirrors the kernel's wait / arrive / issue order; asynchronous completions (TMA
loads, tensor-core commits in issue order, bulk-store reads in issue order) fire after random delays.  The models
assert that every agent terminates (no deadlock, no lost phase) and that no buffer is read before it is produced or
overwritten before it is consumed.  They check the protocol as designed, not the PTX."""
import pytest

from pipeline_sim import Bar, Sim


def _xattn_protocol(ntiles, seed):
    sim = Sim(seed)
    kv_full = Bar(1); q_full = [Bar(1), Bar(1)]; q_empty = [Bar(1), Bar(1)]; s_full = [Bar(1), Bar(1)]
    p_full = [Bar(4), Bar(4)]; o_full = [Bar(1), Bar(1)]; o_empty = [Bar(4), Bar(4)]
    # data versions
    Q = [None, None]; S = [None, None]; P = [None, None]; O = [None, None]
    s_read = [set(), set()]; o_read = [set(), set()]
    kv = [False]
    def tma():
        kv_full.arrive(tx=100); sim.later(lambda: (kv.__setitem__(0, True), kv_full.complete_tx(100)))
        yield ('step',)
        for i in range(ntiles):
            b = i & 1
            yield ('wait', q_empty[b], ((i >> 1) & 1) ^ 1)
            q_full[b].arrive(tx=10)
            def land(b=b, i=i):
                Q[b] = i; q_full[b].complete_tx(10)
            sim.later(land)
            yield ('step',)
    def mma():
        def issue_qk(b, i):
            def done(b=b, i=i):
                assert Q[b] == i, f"QK({i}) read Q[{b}]={Q[b]}"
                # S[b]/P[b] of tile i-2 must have been consumed by PV(i-2)
                assert S[b] is None or S[b][1] == 'consumed', f"QK({i}) overwrites live S[{b}]={S[b]}"
                S[b] = [i, 'fresh']; s_read[b] = set()
            sim.later(done, 'tc')
        def commit(bar):
            sim.later(lambda: bar.arrive(), 'tc')
        yield ('wait', kv_full, 0)
        yield ('wait', q_full[0], 0)
        issue_qk(0, 0); commit(s_full[0])
        yield ('step',)
        for i in range(ntiles):
            b = i & 1
            if i + 1 < ntiles:
                yield ('wait', q_full[b ^ 1], ((i + 1) >> 1) & 1)
                issue_qk(b ^ 1, i + 1); commit(s_full[b ^ 1])
                yield ('step',)
            yield ('wait', p_full[b], (i >> 1) & 1)
            yield ('wait', o_empty[b], ((i >> 1) & 1) ^ 1)
            def pv(b=b, i=i):
                assert P[b] == i, f"PV({i}) read P[{b}]={P[b]}"
                assert O[b] is None or len(o_read[b]) == 4, f"PV({i}) overwrites unread O[{b}]"
                O[b] = i; o_read[b] = set(); S[b][1] = 'consumed'
            sim.later(pv, 'tc')
            commit(o_full[b]); commit(q_empty[b])
            yield ('step',)
    def softmax(w):
        for i in range(ntiles):
            b = i & 1; ph = (i >> 1) & 1
            yield ('wait', s_full[b], ph)
            assert S[b] is not None and S[b][0] == i, f"softmax w{w} tile {i} read S[{b}]={S[b]}"
            s_read[b].add(w)
            yield ('step',)
            # all 4 warps write their quarter of P; mark P[b]=i when the 4th has written (p_full count 4 handles sync)
            if len(s_read[b]) == 4: P[b] = i
            p_full[b].arrive()
            yield ('wait', o_full[b], ph)
            assert O[b] == i, f"epilogue w{w} tile {i} read O[{b}]={O[b]}"
            o_read[b].add(w)
            yield ('step',)
            o_empty[b].arrive()
    sim.add('tma', tma()); sim.add('mma', mma())
    for w in range(4): sim.add(f'sm{w}', softmax(w))
    sim.run()



NB = 3

class NamedBar:
    def __init__(self, n): self.n = n; self.count = 0; self.gen = 0
    def arrive(self):
        self.count += 1
        if self.count == self.n: self.count = 0; self.gen += 1
    def passed(self, gen): return self.gen != gen      # same interface as Bar.passed(parity)

def _tepi_protocol(ntiles, panels_per_group, has_res, seed):
    sim = Sim(seed)
    tfull = [Bar(1), Bar(1)]; tempty = [Bar(8), Bar(8)]
    acc = [None, None]; acc_reads = [0, 0]
    def mma():
        for it in range(ntiles):
            a = it & 1; aph = (it >> 1) & 1
            yield ('wait', tempty[a], aph ^ 1)
            def done(a=a, it=it):
                assert acc[a] is None or acc_reads[a] == 8, f"MMA tile {it} overwrites accumulator {a} still being read ({acc_reads[a]})"
                acc[a] = it; acc_reads[a] = 0
                tfull[a].arrive()
            sim.later(done, 'tc')
            yield ('step',)
    sim.add('mma', mma())
    for grp in range(2):
        my_panels = panels_per_group[grp]
        res_full = [Bar(1) for _ in range(NB)]
        nbar = NamedBar(4)
        # buffer state: dict(state=..., panel=global panel index of this group)
        buf = [dict(state='free', panel=None, writers=0) for _ in range(NB)]
        pending_store_reads = []      # FIFO of buffers with a store that has not finished reading
        pf = dict(q=0, b=0)           # prefetch iterator (global panel counter of this group)
        total_panels = ntiles * my_panels
        def prefetch(buf=buf, res_full=res_full, pf=pf, total_panels=total_panels):
            if pf['q'] >= total_panels: return
            b = pf['b']; q = pf['q']
            assert buf[b]['state'] == 'free', f"grp residual prefetch panel {q} into buffer {b} in state {buf[b]}"
            buf[b].update(state='res_loading', panel=q)
            res_full[b].arrive(tx=1)
            def land(b=b, q=q):
                assert buf[b]['state'] == 'res_loading' and buf[b]['panel'] == q
                buf[b]['state'] = 'res_ready'; res_full[b].complete_tx(1)
            sim.later(land)
            pf['q'] += 1; pf['b'] = (pf['b'] + 1) % NB
        def warp(w, grp=grp, my_panels=my_panels, res_full=res_full, nbar=nbar, buf=buf, pending=pending_store_reads, pf=pf, prefetch=prefetch):
            elected = (w == 0)
            if has_res and elected:
                for _ in range(NB - 1): prefetch()
            yield ('step',)
            b = 0; bphase = 0; q = 0
            for it in range(ntiles):
                a = it & 1; aph = (it >> 1) & 1
                yield ('wait', tfull[a], aph)
                assert acc[a] == it
                for i in range(my_panels):
                    if has_res:
                        yield ('wait', res_full[b], bphase)
                        assert buf[b]['state'] in ('res_ready', 'written') and buf[b]['panel'] == q, f"warp {w} panel {q}: buffer {b} = {buf[b]}"
                    else:
                        assert buf[b]['state'] in ('free', 'written') and (buf[b]['state'] == 'free' or buf[b]['panel'] == q), f"warp {w} panel {q} writes buffer {b} = {buf[b]}"
                    # write my rows
                    if buf[b]['state'] != 'written':
                        buf[b].update(state='written', panel=q, writers=0)
                    buf[b]['writers'] += 1
                    yield ('step',)
                    g = nbar.gen; nbar.arrive()
                    yield ('wait', nbar, g)
                    if elected:
                        assert buf[b]['writers'] == 4, f"store of panel {q} before all warps wrote ({buf[b]['writers']})"
                        buf[b]['state'] = 'storing'; pending.append(b)
                        def read_done(bb=b):
                            assert pending[0] == bb; pending.pop(0); buf[bb].update(state='free', panel=None)
                        sim.later(read_done, ('bulk', grp))
                        # wait_group.read<1>: at most one store still reading
                        class W:   # ad-hoc waitable
                            def passed(self, _): return len(pending) <= 1
                        yield ('wait', W(), 0)
                        if has_res: prefetch()
                        yield ('step',)
                    b += 1; q += 1
                    if b == NB: b = 0; bphase ^= 1
                acc_reads[a] += 1
                tempty[a].arrive()
                yield ('step',)
        for w in range(4):
            sim.add(f'g{grp}w{w}', warp(w))
    sim.run()



@pytest.mark.parametrize("ntiles", [1, 2, 3, 4, 5, 8])
def test_xattn_tc_protocol(ntiles):
    for seed in range(60):
        _xattn_protocol(ntiles, seed)


@pytest.mark.parametrize("has_res", [False, True])
@pytest.mark.parametrize("panels", [(3, 2), (4, 4), (2, 2), (2, 1)])
def test_gemm_tepi_epilogue_protocol(panels, has_res):
    for ntiles in (1, 2, 3, 5):
        for seed in range(25):
            _tepi_protocol(ntiles, panels, has_res, seed)
