#!/bin/bash
# Compile only the float / VEC 4 / z+t instantiations of the two tile-kernel forms (scheme $1, default hybrid = 3) and print static
# SASS statistics of the z loop of each: a few seconds of turnaround for instruction-count work.  Extra nvcc flags: $2...
SCH=${1:-3}; shift
mkdir -p /tmp/probe
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DSCH=$SCH "$@" -I pytv-4d_b200/csrc -cubin -o /tmp/probe/probe.cubin scripts/sass_probe.cu || exit 1
cuobjdump -sass /tmp/probe/probe.cubin > /tmp/probe/probe.sass
cuobjdump -res-usage /tmp/probe/probe.cubin | grep -o "Function [^ ]*\|REG:[0-9]*" | sed 's/Function _ZN5pytvb\([0-9]*\)//' | cut -c1-40 | paste - -
python3 - <<'PY'
import re, collections
fn = None; ins = {}
for line in open('/tmp/probe/probe.sass'):
    m = re.search(r'Function : (\S+)', line)
    if m: fn = m.group(1); ins[fn] = []; continue
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', line)
    if m and fn: ins[fn].append((int(m.group(1), 16), m.group(2).strip()))
for fn, L in ins.items():
    lo = hi = 0
    for a, t in L:
        if re.search(r'\bBRA', t):
            m = re.search(r'(0x[0-9a-f]+)\s*$', t)
            if m:
                tgt = int(m.group(1), 16)
                if tgt < a and a - tgt > hi - lo: lo, hi = tgt, a
    body = [t for a, t in L if lo <= a <= hi]
    c = collections.Counter()
    for t in body:
        t = re.sub(r'^@!?U?P\d+\s+', '', t)
        c[t.split()[0].split('.')[0]] += 1
    name = 'form2' if 'tile2' in fn else 'form1'
    print("%s: %d instructions, z loop %d:  %s" % (name, len(L), len(body), '  '.join('%s %d' % kv for kv in c.most_common(24))))
PY
