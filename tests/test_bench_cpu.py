"""bench.py host-side contract pieces that need no GPU: workload table, clock-sample parsing, and the JSON line of the
`--impl reference` arm (CPU oracle on a reduced sample)."""
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
    monkeypatch.setattr(bench, 'sample_classes', lambda seconds, threads: 1)
    real = bench.cpu_oracle_rate
    monkeypatch.setattr(bench, 'cpu_oracle_rate', lambda n_cls, n_real, threads: real(1, 2, threads))      # 1 class x (2 real + 1 syn) videos
    monkeypatch.setattr(sys, 'argv', ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0'])
    monkeypatch.delenv('RANK', raising=False)
    bench.main()
    lines = [ln for ln in capsys.readouterr().out.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'DM+S2D distill iters/sec' and d['unit'] == 'it/s' and d['higher_is_better']
    assert d['value'] > 0 and d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
    assert d['e2e'] == {'value': d['value'], 'unit': 'it/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['config']['workload'] == bench.WORKLOAD_DESC


def test_other_ranks_of_the_reference_arm_do_nothing(monkeypatch, capsys):
    monkeypatch.setattr(sys, 'argv', ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0'])
    monkeypatch.setenv('RANK', '1')
    bench.main()
    assert capsys.readouterr().out.strip() == ''
