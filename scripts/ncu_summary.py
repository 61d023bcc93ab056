"""Per-launch summary table of an `ncu --page raw --csv` export (units normalised):
   ncu_summary.py file.raw.csv [hbm_peak_GBs]  ->  markdown table on stdout"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6552.6
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
SCALE = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12, 'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3,
         'second': 1, 'usecond': 1e-6, 'msecond': 1e-3, 'nsecond': 1e-9}
def val(r, name):
    if name not in ix or r[ix[name]] == '':
        return float('nan')
    return float(r[ix[name]]) * SCALE.get(units[ix[name]], 1)
TENS = 'sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed'
SMEM = 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'
SMEM2 = 'smsp__inst_executed_pipe_uniform.sum'
print('| # | kernel | grid | time us | tensor pipe % | smem pipe % | DRAM read MB | DRAM write MB | DRAM GB/s | % of HBM peak | regs | smem KB |')
print('|---|---|---|---|---|---|---|---|---|---|---|---|')
for n, r in enumerate(rows[2:]):
    t = val(r, 'gpu__time_duration.sum')
    rd, wr = val(r, 'dram__bytes_read.sum'), val(r, 'dram__bytes_write.sum')
    name = r[ix['Kernel Name']].replace('void ', '').split('(')[0]
    tens = val(r, TENS)
    print(f"| {n} | {name} | {r[ix['launch__grid_size']]} | {t * 1e6:.1f} | {tens:.1f} | {val(r, SMEM):.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | "
          f"{(rd + wr) / t / 1e9:.0f} | {100 * (rd + wr) / t / 1e9 / peak:.1f} | {r[ix['launch__registers_per_thread']]} | "
          f"{val(r, 'launch__shared_mem_per_block_dynamic') / 1e3 if 'launch__shared_mem_per_block_dynamic' in ix else 0:.0f} |")
