"""Liveness / ordering / resource-safety of the ws_gemm_kernel mbarrier protocol (two alternating MMA issuers, loaders,
epilogue) on the executable model in tests/protocol_model.py, for the launch shapes the library uses."""
import pytest

from protocol_model import Launch, simulate

SHAPES = {
    # name: (n_sa, n_sb, n_steps, G, RP, RW, acc_stages, resident)
    'conv0 pair-by-pair': (4, 1, 11, 1, 3, 1, 2, True),
    'conv1': (3, 4, 49, 5, 2, 3, 1, False),
    'conv1 old ring': (3, 4, 49, 7, 2, 2, 1, False),
    'conv2': (49, 2, 12, 6, 2, 2, 1, False),
    'dgrad conv1': (3, 2, 64, 8, 3, 3, 2, False),
    'dgrad conv0': (3, 1, 64, 1, 2, 1, 2, True),
    'wgrad': (5, 1, 8, 8, 2, 2, 2, False),
    'split conv1': (3, 8, 75, 6, 2, 3, 1, False),          # groups of 6 MMAs over ring slots of 4 weight tiles
    'split conv2': (49, 4, 18, 6, 2, 3, 1, False),
    'two-product conv1': (3, 8, 50, 5, 2, 3, 1, False),    # frozen real videos: xh.wh + xh.wl, every weight tile used once
    'two-product conv2': (49, 4, 12, 6, 2, 3, 1, False),
}


@pytest.mark.parametrize('name', sorted(SHAPES))
@pytest.mark.parametrize('n_tiles,grid,cta', [(1, 1, 0), (5, 2, 1), (7, 3, 0), (3, 148, 2), (3, 148, 100)])
def test_classic_generator(name, n_tiles, grid, cta):
    n_sa, n_sb, n_steps, G, RP, RW, acc_stages, resident = SHAPES[name]
    L = Launch(n_tiles, grid, n_sa, n_sb, n_steps, G, RP, RW, acc_stages, resident)
    expected_tiles = len(range(cta, n_tiles, grid))
    groups_per_tile = n_sa * n_sb * (1 if resident else -(-n_steps // G))
    assert simulate(L, cta) == expected_tiles * groups_per_tile


@pytest.mark.parametrize('nu_total,n_u,ug_count', [(4, 4, 1), (7, 4, 2), (74, 19, 4)])
@pytest.mark.parametrize('n_tiles,grid', [(4, 2), (9, 4)])
def test_column_gemm_with_shared_pixel_stage(nu_total, n_u, ug_count, n_tiles, grid):
    # backward column GEMM: one pixel stage per tile reused by n_u accumulator groups (M tiles), last CTA tile partially filled
    L = Launch(n_tiles * ug_count, grid, 1, 1, 8, 8, 2, 2, 2, False, n_u=n_u, nu_total=nu_total, ug_count=ug_count)
    for cta in range(min(grid, 3)):
        tiles = range(cta, n_tiles * ug_count, grid)
        assert simulate(L, cta) == sum(min(n_u, nu_total - (t % ug_count) * n_u) for t in tiles)


@pytest.mark.parametrize('pairs', [2, 4, 6, 8, 16])
@pytest.mark.parametrize('n_tiles,grid,cta', [(1, 1, 0), (4, 2, 1), (5, 3, 0), (2, 148, 1)])
@pytest.mark.parametrize('RP', [2, 3])
def test_conv0_streaming_generator(pairs, n_tiles, grid, cta, RP):
    L = Launch(n_tiles, grid, 2 * pairs, 1, 11, 1, RP, 1, 2, True, stream_pairs=pairs)
    columns = len(range(cta, n_tiles, grid))
    assert simulate(L, cta) == columns * (4 * pairs - 2)           # no MMAs on the two all-zero temporal halo frames


@pytest.mark.parametrize('T', [4, 8, 16, 32])
@pytest.mark.parametrize('n_tiles,grid,cta', [(1, 1, 0), (4, 2, 1), (5, 3, 0), (2, 148, 1)])
@pytest.mark.parametrize('RP', [2, 3, 4])
def test_split_conv0_generator(T, n_tiles, grid, cta, RP):
    # split-fp16 conv 0: 2T stages (frame, part) = one MMA group each, 4 rotating accumulators (output frames)
    L = Launch(n_tiles, grid, T, 2, 11, 1, RP, 1, 4, True, stream_pairs=T, stream_mode=2)
    columns = len(range(cta, n_tiles, grid))
    for latency in (1, 3, 11):
        assert simulate(L, cta, mma_latency=latency) == columns * 2 * T


@pytest.mark.parametrize('T', [4, 8, 16, 32])
@pytest.mark.parametrize('n_tiles,grid,cta', [(1, 1, 0), (4, 2, 1), (5, 3, 0), (2, 148, 1)])
@pytest.mark.parametrize('RP', [2, 3, 4])
def test_two_product_conv0_generator(T, n_tiles, grid, cta, RP):
    # two-product conv 0 (hi part only): T stages = one MMA group each, 4 rotating accumulators
    L = Launch(n_tiles, grid, T, 1, 11, 1, RP, 1, 4, True, stream_pairs=T, stream_mode=2)
    columns = len(range(cta, n_tiles, grid))
    for latency in (1, 3, 11):
        assert simulate(L, cta, mma_latency=latency) == columns * T


@pytest.mark.parametrize('latency', [1, 2, 5, 17])
def test_slow_and_fast_tensor_pipe(latency):
    for name in ('conv1', 'conv0 pair-by-pair'):
        n_sa, n_sb, n_steps, G, RP, RW, acc_stages, resident = SHAPES[name]
        simulate(Launch(4, 2, n_sa, n_sb, n_steps, G, RP, RW, acc_stages, resident), 0, mma_latency=latency)
    simulate(Launch(3, 1, 16, 1, 11, 1, 2, 1, 2, True, stream_pairs=8), 0, mma_latency=latency)


def test_model_detects_a_broken_protocol():
    # sanity of the checker itself: if pix_empty / acc_full expected a third commit that nobody issues, the model must report
    # the deadlock instead of terminating
    import protocol_model as pm
    n_sa, n_sb, n_steps, G, RP, RW, acc_stages, resident = SHAPES['conv0 pair-by-pair']
    L = Launch(3, 1, n_sa, n_sb, n_steps, G, RP, RW, acc_stages, resident)
    init = pm.Barrier.__init__

    def three_commits(self, count):
        init(self, 3 if count == 2 else count)
    pm.Barrier.__init__ = three_commits
    try:
        with pytest.raises(AssertionError, match='deadlock'):
            pm.simulate(L, 0)
    finally:
        pm.Barrier.__init__ = init
