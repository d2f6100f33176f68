"""Static SASS instruction mix of the hot kernels (no GPU needed):

    python profiles/sass_mix.py > profiles/r02c_sass_mix.txt

cuobjdump -sass of multi_agent_solver_b200/libmas_b200.so; per kernel the instruction count by class over the whole
function body.  Static counts, not executed counts -- the loop over the 80 time steps is one copy of its body -- so the
mix is that of the code, which for these kernels is dominated by the time-step loop.  It is the evidence behind
DESIGN.md section 9's "47 % of the instructions are integer / control / load and share the issue port".
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multi_agent_solver_b200", "libmas_b200.so")
KERNELS = [  # (label, regex on the mangled name)
    ("forward_coop_kernel<StLane, 2 step sizes/lane> (line search, full batch)", r"forward_coop_kernelINS_6StLaneELi2ELb0ELb0E"),
    ("forward_coop_kernel<StLane, 1 step size/lane>", r"forward_coop_kernelINS_6StLaneELi1ELb0ELb0E"),
    ("backward_kernel<StLane, example mask 63> (backward pass, full batch)", r"backward_kernelINS_6StLaneELi63E"),
    ("backward_kernel<StLane, all-FD>", r"backward_kernelINS_6StLaneELi0E"),
    ("centralized_kernel<StCirc> (config 5)", r"centralized_kernelINS_6StCircE"),
    ("centralized_mixed_kernel (general stacked solve)", r"centralized_mixed_kernel"),
]
CLASSES = [
    ("fp64 add/mul/fma", r"^(DADD|DMUL|DFMA)"),
    ("fp64 other (DSETP, MUFU.RCP64H, F2F/I2F.F64 ...)", r"^(DSETP|DMNMX|MUFU|F2F|I2F|F2I|FRND)"),
    ("tensor (DMMA)", r"^DMMA"),
    ("global/local load-store (LDG, STG, LDL, STL, LD, ST)", r"^(LDG|STG|LDL|STL|LD|ST|ATOM|RED|CCTL)\b"),
    ("shared / constant loads (LDS, STS, LDC, ULDC)", r"^(LDS|STS|LDC|ULDC|LDSM)"),
    ("integer / logic / move / select", r"^(IADD|IADD3|IMAD|LEA|LOP|LOP3|SHF|SHL|SHR|MOV|SEL|FSEL|ISETP|PLOP3|PRMT|IABS|IMNMX|VIADD|VIMNMX|UIADD3|UMOV|ULOP|ULEA|USHF|UIMAD|USEL|UISETP|R2UR|S2R|S2UR|CS2R|POPC|FLO|BREV|P2R|R2P|SGXT|BMSK|UFLO|UPOPC|UPRMT|R2B|LEPC|VOTE|VOTEU|SHFL|MATCH|REDUX|FADD|FMUL|FFMA|FSETP|FMNMX)"),
    ("control (BRA, BSSY, BSYNC, CALL, RET, EXIT, BAR, WARPSYNC, NOP ...)", r"^(BRA|BRX|BSSY|BSYNC|CALL|RET|EXIT|BAR|WARPSYNC|NOP|YIELD|DEPBAR|BREAK|JMP|ERRBAR|MEMBAR|NANOSLEEP|ACQBULK|ENDCOLLECTIVE|BPT|KILL|PREEXIT)"),
]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)
    print(f"# static SASS instruction mix, {os.path.relpath(LIB, ROOT)} (sm_100a); python profiles/sass_mix.py")
    for label, pat in KERNELS:
        body = next((f for f in funcs if re.match(r"\S*" + pat, f)), None)
        if body is None:
            print(f"\n{label}: not found")
            continue
        ops = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_.]+)?)", body)
        total = len(ops)
        counts = collections.OrderedDict((c, 0) for c, _ in CLASSES)
        other = collections.Counter()
        for op in ops:
            base = op.split(".")[0]
            for c, rx in CLASSES:
                if re.match(rx, base):
                    counts[c] += 1
                    break
            else:
                other[base] += 1
        print(f"\n{label}\n  {body.splitlines()[0][:110]}\n  instructions: {total}")
        for c, n in counts.items():
            if n:
                print(f"    {100.0 * n / total:5.1f} %  {n:6d}  {c}")
        if other:
            n = sum(other.values())
            print(f"    {100.0 * n / total:5.1f} %  {n:6d}  other: " + ", ".join(f"{k} {v}" for k, v in other.most_common(8)))
        d = collections.Counter(op.split(".")[0] for op in ops if re.match(r"^(DADD|DMUL|DFMA)", op))
        print("    fp64 arithmetic: " + ", ".join(f"{k} {v}" for k, v in sorted(d.items())) + "  (-fmad=false: DFMA only where the source asks for fma())")


if __name__ == "__main__":
    sys.exit(main())
