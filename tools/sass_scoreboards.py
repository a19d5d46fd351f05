#!/usr/bin/env python
"""Static check of a cubin / object for the stall this round's profiles kept finding: a consumer that waits on a hardware
scoreboard which a YOUNGER global load also signals.  ptxas tracks every variable-latency result with one of six scoreboards
(write barrier index in the SASS control bits); when several loads of an unrolled ring share one, the first use of the oldest
load's registers also waits for the load issued a few instructions before it (k_dec_unpredict_rgb, ncu r03k: 43 % of the
stall samples on one LOP3).  No GPU needed:

    python tools/sass_scoreboards.py xpng_b200/build/api.o [kernel-name substring] [--window 2000]

For every kernel it decodes stall / write barrier / read barrier / wait mask of each instruction (sm_100a encoding: bits 105..125
of the 128-bit word, as in sm_70+), walks the code linearly and reports the two patterns described at analyse().  LDGSTS
(cp.async) does not count: it signals no register.  Linear walk, so sites right behind a loop's back edge are missed and a site
on a cold path is reported like one in the hot loop: it is a pointer for ncu, not a verdict."""
import re, subprocess, sys
GAP = 48   # instructions between the load that is needed and the younger one that is waited for as well

def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    name, rows = None, []
    for line in out.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name: yield name, rows
            name, rows = m.group(1), []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/", line)
        if m:
            rows.append([int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16), None]); continue
        m = re.match(r"\s+/\* (0x[0-9a-f]{16}) \*/", line)
        if m and rows and rows[-1][3] is None: rows[-1][3] = int(m.group(1), 16)
    if name: yield name, rows

def regs(tok):
    """Registers named by one operand, with the width implied by .64 / descriptors ignored (upper halves are added by the caller)."""
    return [int(r) for r in re.findall(r"\bR(\d+)\b", tok)]

def analyse(rows, window):
    """Two patterns, both per scoreboard b:
       shared   an instruction waits on b, the youngest signal on b still outstanding is a global load that feeds none of its
                operands, and an OLDER global load on b (within `window` instructions) does: it pays for the younger load too;
       serial   a global load that signals b first waits on b while the previous global load on b is outstanding: the loads of
                the ring never overlap."""
    sites, prod = [], {b: [] for b in range(6)}        # barrier -> [index, text, dest regs, outstanding]
    for i, (addr, text, lo, hi) in enumerate(rows):
        if hi is None: continue
        c = hi >> 41
        wr, wait = (c >> 5) & 7, (c >> 11) & 0x3F
        body = re.sub(r"^@!?U?P\d+\s+", "", text)
        op = body.split()[0]
        ops = body[len(op):].split(",")
        is_gload = op.startswith(("LDG", "LD.")) and not op.startswith("LDGSTS")
        srcs = set()
        for t in (ops if op.startswith(("ST", "RED", "ATOM")) else ops[1:]):
            for r in regs(t): srcs.update((r, r + 1))            # .64 address pairs: be generous
        for b in range(6):
            if not (wait >> b & 1): continue
            live = [p for p in prod[b] if p[3]]
            if live:
                young = live[-1]
                if young[4] and not (srcs & young[2]):
                    older = [p for p in prod[b] if p[4] and young[0] - p[0] >= GAP and i - p[0] <= window and (srcs & p[2])]   # neighbours return together: harmless
                    if older: sites.append(("shared", addr, text, b, rows[young[0]][0], young[1]))
                    elif is_gload and wr == b: sites.append(("serial", addr, text, b, rows[young[0]][0], young[1]))
            for p in prod[b]: p[3] = False
        if wr < 6:
            d = set()
            if not op.startswith(("ST", "RED", "LDGSTS", "BAR", "ATOMS")):
                first = regs(ops[0]) if ops else []
                width = 4 if ".128" in op else (2 if ".64" in op else 1)
                for r in first[:1]: d.update(range(r, r + width))
            prod[wr].append([i, body, d, True, is_gload])
            prod[wr] = prod[wr][-64:]
    return sites

if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    window = int(sys.argv[sys.argv.index("--window") + 1]) if "--window" in sys.argv else 2000
    if "--window" in sys.argv: args = [a for a in args if a != str(window)]
    want = args[1] if len(args) > 1 else ""
    for name, rows in kernels(args[0]):
        if want not in name: continue
        sites = analyse(rows, window)
        if sites:
            print(f"== {name}: {len(rows)} instructions, {sum(k[0] == 'shared' for k in sites)} shared / {sum(k[0] == 'serial' for k in sites)} serial site(s)")
            for kind, addr, text, b, ya, yt in sites[:16]:
                print(f"   {kind:6s} {addr:06x}  {text[:64]:64s} waits SB{b}; outstanding on SB{b}: {ya:06x} {yt.split(';')[0][:44]}")
