"""Summarise an `ncu --set full` capture of the small stages (verify, extend, compactions) into
profiles/<tag>_stages_ncu.md.  usage: python tools/stage_profiles.py r01"""
import csv
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = subprocess.run(["ncu", "-i", "gpurun_out/stages_%s.ncu-rep" % tag, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard / issue"),
        ("smsp__inst_executed.sum", "warp instructions"), ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 sectors read")]
seen = set()
with open("profiles/%s_stages_ncu.md" % tag, "w") as f:
    f.write("# %s -- `ncu --set full --clock-control none` of the small stages\n\n" % tag)
    f.write("Command: `python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --slots 1` (10 M pairs per step, 341 k flagged (362 k before the offset sample grid), "
            "13.8 k seeded, 13.7 k anchored reads).  One launch per kernel; times under ncu are serialised and cold, the live "
            "CUDA-event times are in `%s_bench_n1.json` (`roofline.stage_ms_per_step`).\n\n" % tag)
    for r in data:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
        if name in seen:
            continue
        seen.add(name)
        f.write("## `%s`\n\n| metric | value |\n|---|---|\n" % name)
        for key, label in want:
            if key in hdr:
                f.write("| %s (`%s`) | %s %s |\n" % (label, key, r[hdr.index(key)], units[hdr.index(key)]))
        f.write("\n")
print(open("profiles/%s_stages_ncu.md" % tag).read()[:3000])
