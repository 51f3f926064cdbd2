"""Summarise an .ncu-rep (captured on the GPU box with `ncu --set full`) into a small markdown
table for profiles/.  Usage: python tools/summarize_ncu.py gpurun_out/x.ncu-rep profiles/x.md"""
import csv
import io
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
    'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__cycles_elapsed.avg.per_second', 'smsp__cycles_active.avg', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.sum', 'smsp__inst_executed.sum',
]


def main(rep, out):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    units = rows[1]
    data = rows[2:]
    name_i = hdr.index('Kernel Name')
    cols = [(k, hdr.index(k)) for k in KEYS if k in hdr]
    tens = [(h, i) for i, h in enumerate(hdr) if h == 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed']
    with open(out, 'w') as f:
        f.write(f'# ncu summary of `{rep}` (ncu --set full --clock-control none; cold-cache, serialised replays)\n\n')
        for r in data:
            f.write(f'## {r[name_i][:110]}\n\n| metric | value | unit |\n|---|---|---|\n')
            for k, i in cols + [t for t in tens if t not in cols]:
                f.write(f'| {k} | {r[i]} | {units[i]} |\n')
            f.write('\n')
    print('wrote', out, 'kernels:', len(data))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
