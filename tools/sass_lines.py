"""Instructions per source line of one kernel, from the SASS (no GPU needed): the accounting behind DESIGN.md 3.11.

    python tools/sass_lines.py [--kernel SUBSTR] [--lo 0xADDR --hi 0xADDR] [--min N] [extra nvcc flags ...]
Compiles csrc/rg_mpc.cu for sm_100a with -lineinfo, disassembles the cubin with `nvdisasm -g -c`, and prints
  * every backward branch (loop) of the kernel with its instruction / DFMA / LDS counts -- the Cholesky panel loop is the
    one with 84 DFMA, its update loop the one with 32;
  * the instruction count and opcode mix per source line inside [--lo, --hi] (default: the whole kernel).
Default kernel: the lean h = 10 solve kernel (mpc_solve_kernelILi10ELb1)."""
import collections, os, re, subprocess, sys, tempfile
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(REPO, "robot-gym_b200", "csrc")


def main():
    args = sys.argv[1:]
    def opt(name, default):
        if name in args:
            i = args.index(name); v = args[i + 1]; del args[i:i + 2]; return v
        return default
    kernel = opt("--kernel", "mpc_solve_kernelILi10ELb1")
    lo, hi, min_n = int(opt("--lo", "0"), 16), int(opt("--hi", "fffffff"), 16), int(opt("--min", "6"))
    tmp = tempfile.mkdtemp(prefix="sass_lines_")
    obj = os.path.join(tmp, "rg_mpc.o")
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--fmad=true", *args,
                    "-I", os.path.join(REPO, "include"), "-I", CSRC, "-c", os.path.join(CSRC, "rg_mpc.cu"), "-o", obj], check=True)
    subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
    start = [i for i, l in enumerate(dis) if l.startswith(".text.") and kernel in l][0]
    ins, cur = [], None
    for l in dis[start + 1:]:
        if l.startswith("\t.section"): break
        m = re.search(r'//## File ".*?rg_mpc.cu", line (\d+)', l)
        if m: cur = int(m.group(1)); continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m: ins.append((int(m.group(1), 16), cur, m.group(2).strip()))
    print(f"{kernel}: {len(ins)} instructions")
    idx = {a: i for i, (a, _, _) in enumerate(ins)}
    labels = {}
    pos = 0
    for l in dis[start + 1:]:
        if l.startswith("\t.section"): break
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/", l)
        if m: pos = int(m.group(1), 16) + 16
        elif l.startswith(".L_"): labels[l.rstrip(":")] = pos
    for a, _, t in ins:
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?`\((\.L_\w+)\)", t)
        if m and labels.get(m.group(1), 1 << 60) <= a:
            tgt = labels[m.group(1)]
            body = [x for b, _, x in ins if tgt <= b <= a]
            print(f"  loop {tgt:#x}-{a:#x}: {len(body)} instr, {sum('DFMA' in x for x in body)} DFMA, {sum('LDS' in x for x in body)} LDS")
    per, ops = collections.Counter(), collections.defaultdict(collections.Counter)
    for a, ln, t in ins:
        if ln and lo <= a <= hi:
            per[ln] += 1
            ops[ln][(t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0]] += 1
    src = open(os.path.join(CSRC, "rg_mpc.cu")).read().split("\n")
    for ln in sorted(per):
        if per[ln] >= min_n:
            print(f"{ln:5d} {per[ln]:4d}  {src[ln - 1].strip()[:96]:96s} {dict(ops[ln].most_common(4))}")
    print("instructions in range:", sum(per.values()))


if __name__ == "__main__":
    main()
