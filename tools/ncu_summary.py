"""Key metrics per kernel from an ncu report (the raw page): python tools/ncu_summary.py rep [out.json]"""
import csv, io, json, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_uniform.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'lts__t_sector_hit_rate.pct', 'sm__inst_executed.sum',
        'lts__t_sectors_srcunit_tex_op_red.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum',
        'smsp__inst_executed_op_global_red.sum']
res = {}
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')].split('(')[0].replace('void ', '').strip()
    d = {}
    for w in want:
        if w in hdr:
            v = r[hdr.index(w)].replace(',', '')
            try: d[w] = float(v)
            except ValueError: d[w] = v
            d[w + '#unit'] = units[hdr.index(w)]
    res.setdefault(name, []).append(d)
def to_bytes(v, u):
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
traffic = {}
for name, lst in res.items():
    d = max(lst, key=lambda x: x.get('gpu__time_duration.sum', 0))
    tr = to_bytes(d.get('dram__bytes_read.sum', 0), d.get('dram__bytes_read.sum#unit')) + to_bytes(d.get('dram__bytes_write.sum', 0), d.get('dram__bytes_write.sum#unit'))
    traffic[__import__('re').sub(r'<.*>', '', name)] = tr
    print(f"== {name}  ({len(lst)} launches captured)")
    for w in want:
        if w in d: print(f"   {w:72s} {d[w]} {d[w + '#unit']}")
    print(f"   dram traffic per launch: {tr / 1e6:.2f} MB")
if len(sys.argv) > 2:
    json.dump({"source": rep, "dram_bytes_per_launch": traffic}, open(sys.argv[2], "w"), indent=1)
