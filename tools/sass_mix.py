"""Inner-loop instruction mix of the pair kernels, read from the SASS of the built library (no GPU needed).

    make -C cvortex_b200/csrc sass && python tools/sass_mix.py [filter]

For every `m2m_kernel<Policy, T, B, MINB>` it finds the innermost loop (the backward branch whose body
holds the most MUFU instructions per byte), counts opcodes in it, and prints them per (source, target)
pair: packed FP32x2 instructions count two lane-ops, so `lane-ops/pair` is directly comparable with the
L column of DESIGN.md section 4, `MUFU/pair` with S, and `issue/pair` (all instructions) with the
issue bound of 4 warp instructions per clock and SM.
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

SASS = Path(__file__).resolve().parent.parent / "cvortex_b200" / "lib" / "libcvortex.sass"
INS = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*(.*?);")


def functions(text):
	name, body = None, []
	for line in text.splitlines():
		m = re.search(r"Function : (\S+)", line)
		if m:
			if name:
				yield name, body
			name, body = m.group(1), []
			continue
		m = INS.match(line)
		if m and name:
			body.append((int(m.group(1), 16), m.group(2), m.group(3)))
	if name:
		yield name, body


def inner_loop(body):
	"""The backward branch whose span is smallest among those that contain MUFU instructions."""
	best = None
	for k, (addr, op, args) in enumerate(body):
		if not op.startswith("BRA"):
			continue
		m = re.search(r"0x([0-9a-f]+)", args)
		if not m:
			continue
		tgt = int(m.group(1), 16)
		if tgt >= addr:
			continue
		span = [i for i in body if tgt <= i[0] <= addr]
		if not any(i[1].startswith("MUFU") for i in span):
			continue
		if best is None or len(span) < len(best):
			best = span
	return best or []


def main():
	flt = sys.argv[1] if len(sys.argv) > 1 else ""
	text = SASS.read_text()
	names = [n for n, _ in functions(text)]
	dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
	print(f"{'kernel':44s} {'pairs':>5s} {'lane-ops':>8s} {'MUFU':>5s} {'LDS':>5s} {'other':>6s} {'issue':>6s}   (per pair)   top non-FP32 opcodes")
	for (name, body), d in zip(functions(text), dem):
		m = re.search(r"m2m_kernel<cvtx::(\w+(?:<\d+>)?), (\d+), (\d+), (\d+)>", d)
		if not m or flt not in d:
			continue
		pol, T = m.group(1), int(m.group(2))
		loop = inner_loop(body)
		if not loop:
			continue
		c = collections.Counter(op for _, op, _ in loop)
		mufu = sum(v for k, v in c.items() if k.startswith("MUFU"))
		lds = sum(v for k, v in c.items() if k.startswith("LDS"))
		packed = sum(v for k, v in c.items() if re.match(r"F(FMA|MUL|ADD)2", k))
		scalar = sum(v for k, v in c.items() if re.match(r"F(FMA|MUL|ADD)(\.|$)", k) and not re.match(r"F(FMA|MUL|ADD)2", k))
		# sources per trip: one LDS.128 per record
		per_src = 2 if any(p in pol for p in ("P3DVel", "P3DDvort", "P3DVort", "P3DVisc", "F3D")) else 1
		srcs = max(lds // per_src, 1)
		pairs = srcs * T
		lane = packed * 2 + scalar
		total = len(loop)
		other = total - packed - scalar - mufu - lds
		rest = collections.Counter({k: v for k, v in c.items() if not re.match(r"F(FMA|MUL|ADD)", k) and not k.startswith(("MUFU", "LDS"))})
		tops = ", ".join(f"{k}x{v}" for k, v in rest.most_common(5))
		print(f"{pol + f' T={T}':44s} {pairs:5d} {lane / pairs:8.2f} {mufu / pairs:5.2f} {lds / pairs:5.2f} {other / pairs:6.2f} {total / pairs:6.2f}   {tops}")


if __name__ == "__main__":
	main()
