"""Static look at a kernel's SASS: total instructions and the size/mix of its largest backward-branch loop."""
import re, subprocess, sys, collections
lib, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(['cuobjdump','-sass',lib],capture_output=True,text=True).stdout
for p in re.split(r'\n\s*Function : ', txt):
    name = p.split('\n',1)[0]
    if pat not in name: continue
    ins = []
    for l in p.split('\n'):
        m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
        if m: ins.append((int(m.group(1),16), m.group(2)))
    # backward branches
    best = None
    for a, s in ins:
        m = re.search(r'BRA(?:\.\w+)*\s+(?:!?U?P\d+,\s*)?`?\(?\.?L?_?x?_?(\w+)\)?', s)
        m2 = re.search(r'0x([0-9a-f]+)\s*$', s)
        if s.split()[0].startswith('BRA') or (s.startswith('@') and 'BRA' in s):
            if m2:
                tgt = int(m2.group(1),16)
                if tgt < a and (best is None or a - tgt > best[1] - best[0]): best = (tgt, a)
    print(name[:70], 'total', len(ins))
    if best:
        body = [s for a,s in ins if best[0] <= a <= best[1]]
        ops = collections.Counter()
        for s in body:
            t = s.split()
            op = t[1] if t[0].startswith('@') else t[0]
            ops[op.split('.')[0]] += 1
        print('  largest loop', len(body), 'instrs;', dict(ops.most_common(14)))
