#!/bin/bash
# tools/sass_evidence.sh > profiles/rN_sass_evidence.txt -- SASS mnemonic evidence of the built library (no GPU needed)
SO=dynavsr_b200/libdvsr_b200.so
echo "SASS evidence, cuobjdump -sass $SO (sm_100a), mnemonic counts:"
cuobjdump -sass $SO 2>/dev/null | grep -oE "UTCHMMA|UTMALDG\.[0-9]D|UTMAPF\.L2\.[0-9]D|LDTM\.|UTCBAR|UTCATOMSWS\.[A-Z_.]+|SYNCS\.[A-Z0-9_.]+|LDG\.E\.ENL2\.256[A-Z.]*|STG\.E\.ENL2\.256|F2FP\.BF16\.F32\.PACK_AB|FFMA2|RED\.E\.ADD\.F32[A-Z0-9.]*|SHFL\.BFLY|LDS\.128|STS\.128" | sort | uniq -c | sort -rn
echo
echo "UTCHMMA = tcgen05.mma, UTMALDG = cp.async.bulk.tensor (TMA load), UTMAPF = TMA L2 prefetch, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,"
echo "UTCATOMSWS = TMEM allocation, SYNCS = mbarrier ops, LDG/STG.E.ENL2.256 = 256-bit global loads / stores, F2FP.BF16.F32.PACK_AB = packed bf16 split,"
echo "FFMA2 = packed fp32x2 math (DCN blend / split), SHFL.BFLY = quad transposes of the coalesced epilogues, RED.E.ADD.F32 = gradient scatter accumulation."
echo
echo "kernels containing UTCHMMA (count per kernel):"
cuobjdump -sass $SO 2>/dev/null | awk '/Function : /{f=$3} /UTCHMMA/{c[f]++} END{for(k in c) print c[k], k}' | sort -rn
echo
echo "kernels containing UTMALDG (TMA loads):"
cuobjdump -sass $SO 2>/dev/null | awk '/Function : /{f=$3} /UTMALDG/{c[f]++} END{for(k in c) print c[k], k}' | sort -rn
