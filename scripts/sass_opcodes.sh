#!/bin/bash
# Static SASS evidence of the C3 kernel: opcode histogram plus the lines that show the TMA bulk copy (UBLKCP), its
# mbarrier (SYNCS), the read-only costmap loads (LDG.E.CONSTANT), the MUFU uses and the shared-memory exchange
# (LDS.128).  usage: scripts/sass_opcodes.sh > profiles/solve_kernel_r2_sass_opcodes.txt
cd "$(dirname "$0")/.."
OBJ=neo_mpc_planner2_b200/csrc/build/solve_g5.o
FUN=$(cuobjdump -sass $OBJ | grep "Function :" | grep "solve_kernelILi5ELi2ELb0ELb1" | sed 's/.*Function : //')
cuobjdump -sass -fun "$FUN" $OBJ | grep -E "^\s+/\*[0-9a-f]{4}\*/" > /tmp/_sass.txt
N=$(wc -l < /tmp/_sass.txt)
echo "Static SASS of solve_kernel<5,2,false,true> (the C3 kernel: reference fast path, full horizon), sm_100a: cuobjdump -sass $OBJ"
echo "$N instructions.  Opcode histogram (static):"
sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//; s/^@!?U?P[0-9T]+\s+//' /tmp/_sass.txt | awk '{print $1}' | sed 's/\..*//; s/;//' | sort | uniq -c | sort -rn |
  awk -v n=$N '{printf "  %-12s %5d  %5.1f %%\n", $2, $1, 100*$1/n}'
echo
echo "TMA / mbarrier / read-only loads / MUFU / vector shared-memory loads of the exchange:"
grep -E "UBLKCP|SYNCS|LDG\.E\.CONSTANT|MUFU|LDS\.128|LDS\.64|STS\.128|STS\.64" /tmp/_sass.txt | sed -E 's/\s+\/\* 0x[0-9a-f]+ \*\///'
