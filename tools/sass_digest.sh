#!/bin/bash
# SASS digest of the built objects (no GPU needed): per kernel, how many tensor-core (DMMA), TMA
# (UTMALDG), mbarrier (SYNCS), cp.async (LDGSTS), local-memory (LDL/STL), FP64 FMA and MUFU
# instructions it holds.      bash tools/sass_digest.sh > profiles/r02_sass_digest.txt
cd "$(dirname "$0")/.."
printf "%6s %8s %6s %7s %5s %5s %6s %5s  %s\n" DMMA UTMALDG SYNCS LDGSTS LDL STL DFMA MUFU kernel
for o in build/obj/gemm.o build/obj/potrf.o build/obj/gram.o build/obj/gpr.o build/obj/handle.o build/obj/adjoint.o; do
  [ -f "$o" ] || continue
  echo "== $o"
  cuobjdump -sass "$o" | awk '
    function flush() { if (name != "") printf "%6d %8d %6d %7d %5d %5d %6d %5d  %s\n", dmma, tma, syncs, ldgsts, ldl, stl, dfma, mufu, name }
    /Function :/ { flush(); name=$3; dmma=tma=syncs=ldgsts=ldl=stl=dfma=mufu=0 }
    /DMMA/ {dmma++} /UTMALDG/ {tma++} /SYNCS/ {syncs++} /LDGSTS/ {ldgsts++} /LDL/ {ldl++} /STL/ {stl++} /DFMA/ {dfma++} /MUFU/ {mufu++}
    END { flush() }' | c++filt 2>/dev/null | sed -E 's/\(anonymous namespace\):://g; s/\(.*$//'
done
