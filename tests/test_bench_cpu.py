"""bench.py host-side contract pieces that need no GPU: workload table, clock-sample parsing, and the JSON line of the
`--impl reference` arm (the reference's own modules on a bounded sample)."""
import json
import sys

import pytest

import bench


def test_workload_table_matches_baseline_configs():
    bench.set_workload('U-ipc1')
    assert (bench.C, bench.T, bench.HW, bench.VPC, bench.SPC, bench.DPC, bench.BATCH_REAL) == (50, 16, 112, 1, 2, 2, 64)
    assert 'configs[1]' in bench.WORKLOAD_DESC and '50 classes' in bench.WORKLOAD_DESC
    # algorithmic FLOP per embedded video: 11.002 / 1.773 GFLOP
    assert abs(bench.F_EMBED - 11.002e9) < 2e6
    bench.set_workload('K-ipc5')
    assert (bench.C, bench.T, bench.HW, bench.VPC) == (400, 8, 64, 5) and abs(bench.F_EMBED - 1.772e9) < 2e6
    bench.set_workload('U-ipc1')


def test_clock_sampler_parses_and_windows_samples():
    s = bench.ClockSampler(0)
    s.proc = object()                      # pretend nvidia-smi ran
    s.proc = type('P', (), {'terminate': lambda self: None, 'wait': lambda self, timeout=None: 0, 'kill': lambda self: None})()
    rows = [['0', '1965', '1965', '300.0', 'Active', 'Not Active', 'Not Active', 'Not Active', 'Not Active', 10.0],
            ['0', '1750', '1965', '990.5', 'Active', 'Not Active', 'Not Active', 'Not Active', 'Active', 20.0],
            ['0', '1800', '1965', '980.0', 'Active', 'Not Active', 'Not Active', 'Not Active', 'Active', 21.0],
            ['0', '1200', '1965', '200.0', 'Active', 'Active', 'Not Active', 'Not Active', 'Not Active', 40.0]]
    s.rows = rows
    s.mark(19.5, 21.5)                     # only the two samples inside the timed window count
    out = s.stop()
    assert out['sm_mhz'] == 1775.0 and out['sm_max_mhz'] == 1965.0 and out['samples'] == 2
    assert out['reasons'] == ['sw_power_cap'] and out['power_w_max'] == 990.5


def test_reference_arm_prints_one_contract_line(monkeypatch, capsys):
    # a tiny stand-in for the workload (T = 4 frames, 2 real videos per class) so that the reference's own modules finish in seconds
    monkeypatch.setattr(bench, 'T', 4)
    monkeypatch.setattr(bench, 'BATCH_REAL', 2)
    monkeypatch.setattr(bench, 'C', 10)
    monkeypatch.setattr(bench, 'sample_classes', lambda seconds, threads: 1)
    real = bench.cpu_oracle_rate
    monkeypatch.setattr(bench, 'cpu_oracle_rate', lambda n_cls, n_real, threads: real(1, 2, threads))
    monkeypatch.setattr(sys, 'argv', ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0'])
    monkeypatch.delenv('RANK', raising=False)
    monkeypatch.setattr(bench, 'set_workload', lambda name: None)
    bench.main()
    lines = [ln for ln in capsys.readouterr().out.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'DM+S2D distill iters/sec' and d['unit'] == 'it/s' and d['higher_is_better']
    assert d['value'] > 0 and d['cpu_baseline']['cores'] >= 1
    # the reference's own modules when they are reachable (build container: /root/reference, GPU box: baseline/_ref), else the port
    assert d['cpu_baseline']['kind'] == ('reference' if bench.reference_dir() else 'port')
    if bench.reference_dir():
        assert '10 of 10 classes' in d['cpu_baseline']['sample']          # at least 10 classes are actually run
        # a step of this arm is the bounded sample it actually ran: steps x ms_per_step is real elapsed time
        assert abs(d['ms_per_step'] - 1000.0 * d['cpu_baseline']['sample_seconds']) < 1e-6
    assert d['e2e'] == {'value': d['value'], 'unit': 'it/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['config'] == bench.bench_config(1) and 'implementation' not in d['config']


def test_both_arms_share_one_config(monkeypatch):
    bench.set_workload('U-ipc1')
    c1, c8 = bench.bench_config(1), bench.bench_config(8)
    assert c1['workload'] == bench.WORKLOAD_DESC and set(c1) == {'workload', 'parallelism', 'real_videos_per_step', 'syn_videos_per_step', 'l2_policy'}
    assert c1['parallelism'] == 'single GPU' and 'c%8' in c8['parallelism']


def test_other_ranks_of_the_reference_arm_do_nothing(monkeypatch, capsys):
    monkeypatch.setattr(sys, 'argv', ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0'])
    monkeypatch.setenv('RANK', '1')
    bench.main()
    assert capsys.readouterr().out.strip() == ''
