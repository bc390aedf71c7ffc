"""Condense an .ncu-rep into the per-launch summary kept under profiles/ (run here: ncu reads reports without a GPU).

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_name.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('dram__bytes_read.sum', 'dram read'),
    ('dram__bytes_write.sum', 'dram write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram % of peak'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 % of peak'),
    ('sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'tensor pipe active % (elapsed)'),
    ('sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active', 'hmma inst % (active)'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM % of peak'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue active %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
    ('sm__cycles_elapsed.avg.per_second', 'SM clock'),
    ('launch__registers_per_thread', 'registers/thread'),
    ('launch__shared_mem_per_block_dynamic', 'dynamic smem/block'),
    ('l1tex__m_xbar2l1tex_read_bytes.sum', 'L2->SM read bytes'),
    ('l1tex__m_l1tex2xbar_write_bytes.sum', 'SM->L2 write bytes'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem bank conflicts'),
]


def main(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f'# {path}: {len(rows) - 2} profiled launches (ncu --set full --clock-control none; times are cold-cache, serialised)')
    for r in rows[2:]:
        print(f'\n== {r[col["Kernel Name"]][:100]}  grid {r[col["Grid Size"]]} block {r[col["Block Size"]]}')
        for key, label in KEYS:
            hits = [h for h in hdr if h == key or h.endswith('.' + key)]
            for h in hits[:1]:
                print(f'   {label:34s} {r[col[h]]} {units[col[h]]}')


if __name__ == '__main__':
    main(sys.argv[1])
