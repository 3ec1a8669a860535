"""ncu raw-page CSV of the A-streaming kernels -> a compact summary CSV and profiles/ncu_traffic.json.

    ncu -i gpurun_out/r02_full_65536.ncu-rep --page raw --csv > gpurun_out/r02_full_65536_raw.csv
    python tools/ncu_summary.py gpurun_out/r02_full_65536_raw.csv profiles/r02_ncu_full_65536_k32_summary.csv profiles/ncu_traffic.json
"""
import csv
import json
import sys

KEEP = ['ID', 'Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'gpu__time_duration.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum',
        'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__shared_mem_per_block_dynamic', 'sm__cycles_active.avg',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'lts__t_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum']


def to_bytes(val, unit):
    mult = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}.get(unit, 1.0)
    return float(val.replace(',', '')) * mult


def main(raw, out_csv, out_json):
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    with open(out_csv, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])
    col = {k: hdr.index(k) for k in ('Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum')}
    names = {'tc_pass_kernel<32, 0>': 'fro:ah', 'tc_pass_kernel<32, 1>': 'fro:wta', 'tc_kl_kernel<0>': 'kl:kl_uht', 'tc_kl_kernel<1>': 'kl:kl_wtu',
             'tc_pass_kernel<(int)32, (int)0>': 'fro:ah', 'tc_pass_kernel<(int)32, (int)1>': 'fro:wta',
             'tc_kl_kernel<(int)0>': 'kl:kl_uht', 'tc_kl_kernel<(int)1>': 'kl:kl_wtu'}
    kernels = {}
    for r in rows[2:]:
        for pat, key in names.items():
            if pat in r[col['Kernel Name']]:
                rd = to_bytes(r[col['dram__bytes_read.sum']], units[col['dram__bytes_read.sum']])
                wr = to_bytes(r[col['dram__bytes_write.sum']], units[col['dram__bytes_write.sum']])
                kernels[key] = {'dram_bytes_per_launch': rd + wr, 'ncu_ms': float(r[col['gpu__time_duration.sum']])}
    json.dump({'source': 'ncu --set full --clock-control none, tools/prof_tc.py --m 65536 --n 65536 --k 32 --kl (%s)' % out_csv,
               'm_loc': 65536, 'n': 65536, 'k': 32, 'dtype': 'f32', 'kernels': kernels}, open(out_json, 'w'), indent=1)
    print(json.dumps(kernels))


if __name__ == '__main__':
    main(*sys.argv[1:4])
