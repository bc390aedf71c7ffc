"""Executable model of the mbarrier protocol of ws_gemm_kernel (video_distillation_b200/csrc/tc_conv.cu), test infrastructure.

Five kinds of actors share a CTA: the pixel loader, the weight loader, TWO MMA issuer threads that walk the same generator of
MMA groups and issue alternate groups (baton hand-over), and the epilogue.  They only synchronise through mbarriers with phase
parity:  pix_full / pix_empty [RP], w_full / w_empty [RW], acc_full / acc_empty [acc_stages], baton [2].  tcgen05.commit makes an
mbarrier arrive when every MMA previously issued BY THAT THREAD has completed; pix_empty and acc_full therefore expect one commit
from each issuer.  The model replays the protocol with a cooperative scheduler and an in-order tensor pipe and checks

  * liveness : every actor terminates (no deadlock) for the launch shapes the library uses;
  * order    : MMA groups enter the pipe in exactly the sequential generator order (bitwise reproducible accumulation);
  * safety   : a pixel / weight slot is never refilled while a group that reads it is still in flight, an accumulator is never
               re-initialised before the epilogue drained it, and the epilogue never drains an accumulator with MMAs in flight.

It mirrors the control flow of the kernel (`next` / `next_stream` generators, loader and epilogue loops); the arithmetic is not
modelled.  `python -m pytest tests/test_protocol_model_cpu.py`.
"""
from collections import deque


class Barrier:
    """mbarrier with an arrival count; wait(parity) passes once the phase with that parity has completed."""

    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, 'more arrivals than the barrier expects in one phase'
        if self.pending == 0:
            self.pending = self.count
            self.phase ^= 1

    def passed(self, parity):
        return self.phase != parity


class Launch:
    """Launch shape (the fields of WsParams the control flow depends on)."""

    def __init__(self, n_tiles, grid, n_sa, n_sb, n_steps, G, RP, RW, acc_stages, resident, n_u=1, nu_total=1, ug_count=1,
                 stream_pairs=0, stream_mode=0):
        self.__dict__.update(locals())
        self.G = n_steps if resident else G


def classic_groups(L, cta):
    """tc_conv.cu `next`: (tile, u, stage, weight-slot group) -> group record."""
    pslot = pphase = wslot = wphase = as_ = aphase = 0
    spt = L.n_sa * L.n_sb
    slots_per_stage = (L.n_steps + L.G - 1) // L.G
    tile = cta
    while tile < L.n_tiles:
        n_u_eff = min(L.n_u, L.nu_total - (tile % L.ug_count) * L.n_u)
        for u in range(n_u_eff):
            for st in range(spt):
                for g in range(slots_per_stage):
                    first, last_g, last_u = (st == 0 and g == 0), g == slots_per_stage - 1, u == n_u_eff - 1
                    yield dict(w_acc=('acc_empty', as_, aphase ^ 1), w_pix=('pix_full', pslot, pphase),
                               w_w=None if L.resident else ('w_full', wslot, wphase),
                               acc=as_, first=first, pix=pslot, wt=None if L.resident else wslot,
                               c_w=None if L.resident else ('w_empty', wslot),
                               c_pix=('pix_empty', pslot) if (last_g and last_u) else None,
                               c_acc=('acc_full', as_) if (last_g and st == spt - 1) else None)
                    if not L.resident:
                        wslot += 1
                        if wslot == L.RW:
                            wslot, wphase = 0, wphase ^ 1
                if last_u:
                    pslot += 1
                    if pslot == L.RP:
                        pslot, pphase = 0, pphase ^ 1
            as_ += 1
            if as_ == L.acc_stages:
                as_, aphase = 0, aphase ^ 1
        tile += L.grid


def stream_groups(L, cta):
    """tc_conv.cu `next_stream` (conv 0, input-frame streaming)."""
    pslot = pphase = 0
    qbase = 0
    pairs = L.stream_pairs
    tile = cta
    while tile < L.n_tiles:
        for fi in range(2 * pairs):
            pp, odd = divmod(fi, 2)
            kinds = ([0] if pp > 0 else []) + [1] if not odd else [2] + ([3] if pp + 1 < pairs else [])
            for kind in kinds:
                q = qbase + pp + (-1 if kind == 0 else 1 if kind == 3 else 0)
                buf = q & 1
                first_write = kind == 3 or (kind == 1 and pp == 0)
                final_write = kind == 0 or (kind == 2 and pp == pairs - 1)
                last_of_frame = kind == kinds[-1]
                yield dict(w_acc=('acc_empty', buf, ((q >> 1) & 1) ^ 1) if first_write else None, w_pix=('pix_full', pslot, pphase),
                           w_w=None, acc=buf, first=first_write, pix=pslot, wt=None, c_w=None,
                           c_pix=('pix_empty', pslot) if last_of_frame else None, c_acc=('acc_full', buf) if final_write else None)
            pslot += 1
            if pslot == L.RP:
                pslot, pphase = 0, pphase ^ 1
        qbase += pairs
        tile += L.grid


def l0s_groups(L, cta):
    """tc_conv.cu `next_l0s` (split-fp16 conv 0): ONE group per stage (frame i, part) with up to three segments, kt = 2, 1, 0 ->
    output frame i + 1 - kt; frame f of the running count lives in accumulator buffer (qbase + f) & 3."""
    pslot = pphase = 0
    qbase = 0
    T = L.stream_pairs
    tile = cta
    while tile < L.n_tiles:
        nparts = L.n_sb                         # 2: (frame, hi / lo part) stages; 1: the two-product mode (hi part only)
        for ss in range(nparts * T):
            i, part = divmod(ss, nparts)
            segs, w_accs, c_accs = [], [], []
            for kt in (2, 1, 0):
                f = i + 1 - kt
                if not (0 <= f < T):
                    continue
                q = qbase + f
                buf = q & 3
                first_write = part == 0 and (kt == 0 or (f == 0 and kt == 1))
                final_write = part == nparts - 1 and (kt == 2 or (f == T - 1 and kt == 1))
                segs.append((buf, first_write))
                if first_write:
                    w_accs.append(('acc_empty', buf, ((q >> 2) & 1) ^ 1))
                if final_write:
                    c_accs.append(('acc_full', buf))
            assert len(w_accs) <= 2 and len(c_accs) <= 2
            yield dict(w_acc=w_accs, w_pix=('pix_full', pslot, pphase), w_w=None, segs=segs, pix=pslot, wt=None, c_w=None,
                       c_pix=('pix_empty', pslot), c_acc=c_accs, nmma=len(segs))
            pslot += 1
            if pslot == L.RP:
                pslot, pphase = 0, pphase ^ 1
        qbase += T
        tile += L.grid


def _norm(g):
    """Group records of the single-accumulator generators in the multi-segment form used by simulate()."""
    if 'segs' not in g:
        g['segs'] = [(g['acc'], g['first'])]
        g['w_acc'] = [g['w_acc']] if g['w_acc'] is not None else []
        g['c_acc'] = [g['c_acc']] if g['c_acc'] is not None else []
    return g


def simulate(L, cta=0, mma_latency=3):
    """Runs one CTA of the launch; returns the number of MMA groups issued.  Raises AssertionError on a violation."""
    bars = {('pix_full', i): Barrier(1) for i in range(L.RP)}
    bars.update({('pix_empty', i): Barrier(2) for i in range(L.RP)})
    bars.update({('w_full', i): Barrier(1) for i in range(L.RW)})
    bars.update({('w_empty', i): Barrier(1) for i in range(L.RW)})
    bars.update({('acc_full', i): Barrier(2) for i in range(L.acc_stages)})
    bars.update({('acc_empty', i): Barrier(1) for i in range(L.acc_stages)})      # the 4 epilogue warps modelled as one actor
    bars.update({('baton', i): Barrier(1) for i in range(2)})
    groups = [_norm(g) for g in (l0s_groups if L.stream_mode == 2 else stream_groups if L.stream_pairs else classic_groups)(L, cta)]
    pipe = deque()                       # in-flight groups: [remaining time, group index, issuer]
    issued_by = [-1, -1]                 # index of the last group each issuer put into the pipe
    done = -1                            # index of the last completed group
    pending_commits = []                 # (issuer, last group index issued by it at commit time, barrier key)
    in_flight_reads = {'pix': {}, 'wt': {}}        # slot -> set of group indices issued and not completed
    acc_state = {i: 'free' for i in range(L.acc_stages)}      # free -> accumulating -> full -> free
    order = []

    def wait(key_parity):
        key, parity = key_parity[:2], key_parity[2]
        while not bars[key].passed(parity):
            yield

    def pixel_loader():
        slot = phase = 0
        stages = sum(1 for g in groups if g['c_pix'])
        for _ in range(stages):
            yield from wait(('pix_empty', slot, phase ^ 1))
            assert not in_flight_reads['pix'].get(slot), 'pixel slot refilled while MMAs still read it'
            bars[('pix_full', slot)].arrive()
            slot += 1
            if slot == L.RP:
                slot, phase = 0, phase ^ 1

    def weight_loader():
        slot = phase = 0
        for g in groups:
            if g['wt'] is None:
                continue
            yield from wait(('w_empty', slot, phase ^ 1))
            assert not in_flight_reads['wt'].get(slot), 'weight slot refilled while MMAs still read it'
            bars[('w_full', slot)].arrive()
            slot += 1
            if slot == L.RW:
                slot, phase = 0, phase ^ 1

    def issuer(role):
        for k, g in enumerate(groups):
            if (k & 1) == role:
                # (key, index, parity) waits are taken before the group; (key, index, parity, segment) waits — split-fp16 conv 0 —
                # just before the MMAs of that segment, with the earlier segments of the group already in the pipe
                for w in [w for w in g['w_acc'] if len(w) == 3 or w[3] == 0] + [g['w_pix'], g['w_w']]:
                    if w is not None:
                        yield from wait(w)
                if k > 0:
                    yield from wait(('baton', role ^ 1, ((k - 1) >> 1) & 1))
                for si, (acc, first) in enumerate(g['segs']):
                    for w in g['w_acc']:
                        if len(w) == 4 and w[3] == si and si > 0:
                            yield from wait(w)
                    if first:
                        assert acc_state[acc] == 'free', 'accumulator re-initialised before the epilogue drained it'
                        acc_state[acc] = 'accumulating'
                    else:
                        assert acc_state[acc] == 'accumulating', 'accumulate into an accumulator that was not initialised'
                order.append(k)
                pipe.append([mma_latency, k, role])
                issued_by[role] = k
                in_flight_reads['pix'].setdefault(g['pix'], set()).add(k)
                if g['wt'] is not None:
                    in_flight_reads['wt'].setdefault(g['wt'], set()).add(k)
                bars[('baton', role)].arrive()
                if g['c_w']:
                    pending_commits.append((role, issued_by[role], g['c_w']))
            for c in [g['c_pix']] + g['c_acc']:          # both issuers commit: "MY MMAs up to here have completed"
                if c:
                    pending_commits.append((role, issued_by[role], c))
            yield

    def epilogue():
        as_ = aphase = 0
        for gi, g in enumerate(groups):
            for c in g['c_acc']:
                yield from wait(('acc_full', as_, aphase))
                assert c[1] == as_, 'epilogue and issuers disagree on the accumulator stage'
                # MMAs of LATER groups may already be in flight into other accumulators; none may touch this one
                assert acc_state[as_] == 'accumulating' and all(k > gi or all(a != as_ for a, _ in groups[k]['segs']) for _, k, _ in pipe) \
                    and all(k > gi or True for _, k, _ in pipe), 'epilogue drains an accumulator with MMAs in flight'
                assert not any(k <= gi and any(a == as_ for a, _ in groups[k]['segs']) for _, k, _ in pipe), \
                    'epilogue drains an accumulator with MMAs in flight'
                acc_state[as_] = 'free'
                yield                                          # drain time
                bars[('acc_empty', as_)].arrive()
                as_ += 1
                if as_ == L.acc_stages:
                    as_, aphase = 0, aphase ^ 1

    actors = [pixel_loader(), weight_loader(), issuer(0), issuer(1), epilogue()]
    alive = [True] * len(actors)
    idle_rounds = 0
    while any(alive):
        progressed = False
        snapshot = (len(order), done, tuple(b.phase for b in bars.values()), tuple(b.pending for b in bars.values()))
        for i, a in enumerate(actors):
            if alive[i]:
                try:
                    next(a)
                except StopIteration:
                    alive[i] = False
                    progressed = True
        if pipe:                                           # the tensor pipe executes groups in issue order
            pipe[0][0] -= 1
            if pipe[0][0] == 0:
                _, k, _ = pipe.popleft()
                done = k
                for kind in ('pix', 'wt'):
                    for s in in_flight_reads[kind].values():
                        s.discard(k)
            progressed = True
        for c in list(pending_commits):                    # a commit arrives when the issuer's MMAs up to it are complete
            role, upto, key = c
            if upto <= done:
                bars[key].arrive()
                pending_commits.remove(c)
                progressed = True
        after = (len(order), done, tuple(b.phase for b in bars.values()), tuple(b.pending for b in bars.values()))
        idle_rounds = 0 if (progressed or after != snapshot) else idle_rounds + 1
        assert idle_rounds < 4, f'deadlock: {sum(alive)} actors blocked after {len(order)} of {len(groups)} groups'
    assert order == list(range(len(groups))), 'MMA groups entered the pipe out of order'
    assert not pipe and not pending_commits
    return len(groups)
