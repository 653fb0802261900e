#!/bin/bash
# SASS evidence of the vote kernel kept under profiles/ (run HERE: cuobjdump needs no GPU).
#   tools/sass_excerpt.sh > profiles/r02_sass_k_vote.txt
so=fastposecnn_b200/libfpc_b200.so
fn=$(cuobjdump -sass $so 2>/dev/null | grep -o "_ZN3fpc6k_voteILi0ELb1EEE[A-Za-z0-9_]*" | head -1)
cuobjdump -sass -fun "$fn" $so 2>/dev/null > /tmp/k_vote.sass
echo "# cuobjdump -sass of k_vote<FPC_ARITH_IEEE, packed> in $so (sm_100a), $(date -u +%F)"
echo "# resource usage:"; cuobjdump --dump-resource-usage $so 2>/dev/null | grep -A1 "$fn" | tail -1
echo; echo "# mnemonic counts, whole kernel (UBLKCP = cp.async.bulk, SYNCS = mbarrier, FFMA2 = packed f32x2 FMA, LDGSTS = cp.async):"
grep -E "^\s+/\*[0-9a-f]{4}\*/" /tmp/k_vote.sass | sed -E 's/^\s+\/\*[0-9a-f]{4}\*\/\s+//; s/^@!?U?P[0-9T] +//' | awk '{print $1}' | sed 's/\..*//' | sort | uniq -c | sort -rn | awk '{printf "%s %s, ", $2, $1} END {print ""}' | fold -w 150
echo; echo "# full mnemonics of the Blackwell-specific instructions:"
grep -oE "(UBLKCP|SYNCS|LDGSTS|FFMA2|FMNMX3|REDUX|CREDUX|MATCH|ATOMS|LEA\.HI)[A-Z0-9a-z_.]*" /tmp/k_vote.sass | sort | uniq -c | sort -rn
# hot loop = the longest run of lines between two backward branches that is dominated by FFMA2
python3 - <<'PY'
import re
lines=[l.rstrip() for l in open("/tmp/k_vote.sass") if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
addr=lambda l:int(re.match(r"\s+/\*([0-9a-f]{4})\*/", l).group(1),16)
best=None
for i,l in enumerate(lines):
    m=re.search(r"BRA\s+0x([0-9a-f]+)", l)
    if m and int(m.group(1),16) < addr(l):
        t=int(m.group(1),16); body=[x for x in lines if t <= addr(x) <= addr(l)]
        n2=sum("FFMA2" in x for x in body)
        if n2 >= 100 and (best is None or len(body) < len(best)): best=body
print(f"\n# hot loop: one 16-pixel round x 4 hypotheses per lane = 32 pairs of votes, {len(best)} instructions:")
import collections
c=collections.Counter(re.sub(r"\..*","",re.sub(r"^\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P[0-9T]\s+)?","",x).split()[0]) for x in best)
print("#   "+", ".join(f"{k} {v}" for k,v in c.most_common()))
for x in best: print(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$","",x))
PY
