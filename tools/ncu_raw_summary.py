"""Per-kernel summary of an ncu report's raw page.
usage: ncu -i X.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_raw_summary.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "lsu_wb%"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu_wave%"),
        ("lts__t_bytes.sum", "l2_bytes"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("sm__inst_executed.avg.per_cycle_elapsed", "ipc"), ("launch__registers_per_thread", "regs"),
        ("smsp__inst_executed_pipe_xu.sum", "xu"), ("sm__cycles_active.avg", "cycles")]
for r in rows[2:]:
    name = r[ix["Kernel Name"]][:40]
    out = []
    for k, short in want:
        if k in ix:
            out.append("%s=%s%s" % (short, r[ix[k]], units[ix[k]].replace("byte", "B").replace("second", "s")))
    print(name, " ".join(out))
