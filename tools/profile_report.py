"""profiles/ncu_unit_kernels_r02.txt from one `ncu --set full` report of the bench workload (tools/gpu_r2_final_a.sh):
per-kernel summary, then stall samples by code region and the hottest source lines of the walk / moment kernels."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
src = open(os.path.join(ROOT, "celeste.jl_b200", "csrc", "unit_kernels.cuh")).read().splitlines()


def line_of(pattern, nth=0):
    hits = [i + 1 for i, l in enumerate(src) if pattern in l]
    return hits[nth]


walk = line_of("    unit_walk_kernel(")
regions = [f"pixel_term:{line_of('double unit_pixel_term(')}-{line_of('unit_moments_to_sums(') - 1}",
           f"prologue:{walk}-{line_of('phase A: the active') - 1}",
           f"walk_setup:{line_of('phase A: the active')}-{line_of('int t = 0;', 1) - 1}",
           f"walk_start:{line_of('int t = 0;', 1)}-{line_of('int t = 0;', 1) + 8}",
           f"mixture_sums:{line_of('int t = 0;', 1) + 9}-{line_of('float xf = nanf') - 1}",
           f"pixel_loads_star:{line_of('float xf = nanf')}-{line_of('double l5 = 0.0') - 1}",
           f"pixel_call:{line_of('double l5 = 0.0')}-{line_of('fixed-order warp reduction') - 1}",
           f"reduce:{line_of('fixed-order warp reduction')}-{line_of('    unit_moment_kernel(') - 6}",
           f"moment_kernel:{line_of('    unit_moment_kernel(')}-{line_of('Host side: the unit list') - 1}"]
out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
# one summary per distinct kernel (the report holds repeated launches)
seen, keep, cur = set(), [], None
for l in out.splitlines():
    if l.startswith("=="):
        name = re.search(r"ncu-rep (?:void )?(\S+?)\(", l).group(1)
        cur = name not in seen
        seen.add(name)
    if cur:
        keep.append(l)
print("\n".join(keep))
# launch order in the report: slotbr bg walk<1> epi<1> | slotbr bg walk<2> moment epi_hess | ...
for title, skip in (("unit_walk_kernel<1>", 2), ("unit_walk_kernel<2>", 6), ("unit_moment_kernel", 7)):
    env = dict(os.environ, NCU_ARGS=f"--launch-skip {skip} --launch-count 1")
    print(f"\n==== {title}: warp-state samples by code region (percent of the kernel's samples)")
    print(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_regions.py"), rep, "unit_kernels.cuh"] + regions,
                         capture_output=True, text=True, env=env).stdout)
    print(f"==== {title}: hottest source lines")
    print(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "16"], capture_output=True, text=True,
                         env=env).stdout)
