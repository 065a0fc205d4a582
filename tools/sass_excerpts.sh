#!/bin/bash
# SASS evidence of the built library for profiles/: which Blackwell instructions the hot kernels hold, and the inner loops.
# usage: tools/sass_excerpts.sh <tag>      (writes profiles/sass_dense_gemm_<tag>.txt, sass_bm25_score_<tag>.txt, sass_bm25_score16_<tag>.txt)
set -eu
TAG=${1:-r02}
LIB=modern-search-engines-project_b200/csrc/libmsegpu.so
TMP=$(mktemp)
cuobjdump -sass $LIB | grep -v '^\s*/\* 0x' > $TMP
kernel() { awk -v k="$1" '/Function : /{f=(index($0,k)>0)} f' $TMP; }
{
  echo "# cuobjdump -sass $LIB — dense_gemm_kernel<1> (cluster of two CTAs, TMA multicast, cta_group::1 MMAs)"
  echo "# instruction counts:"
  kernel dense_gemm_kernelILi1E | grep -o 'UTCHMMA\|UTMALDG\.2D\(\.MULTICAST\)\?\|LDTM\.x32\|STTM\.x32\|UTCBAR\(\.MULTICAST\)\?\|UTCATOMSWS[A-Z._]*\|SYNCS\.[A-Z0-9.]*\|ELECT\|FMNMX3\|UCGABAR_[A-Z]*\|R2UR' | sort | uniq -c | sort -rn
  echo
  echo "# MMA issue of one k-block (A in tensor memory): four UTCHMMA back to back, then the commit that frees the stage"
  L=$(kernel dense_gemm_kernelILi1E | grep -n 'UTCHMMA' | sed -n 1p | cut -d: -f1)
  kernel dense_gemm_kernelILi1E | sed -n "$((L-34)),$((L+8))p" | cut -c1-118
  echo
  echo "# TMA producer: one half stage = four multicast boxes"
  L=$(kernel dense_gemm_kernelILi1E | grep -n 'UTMALDG.2D.MULTICAST' | sed -n 5p | cut -d: -f1)
  kernel dense_gemm_kernelILi1E | sed -n "$((L-22)),$((L+10))p" | cut -c1-118
  echo
  echo "# epilogue: accumulator -> registers (4 x LDTM.x32), group maximum (FMNMX3), vote"
  L=$(kernel dense_gemm_kernelILi1E | grep -n 'LDTM.x32' | sed -n 1p | cut -d: -f1)
  kernel dense_gemm_kernelILi1E | sed -n "$((L-2)),$((L+50))p" | cut -c1-118
  echo
  echo "# dense_gemm_kernel<2> (cta_group::2 variant) instruction counts:"
  kernel dense_gemm_kernelILi2E | grep -o 'UTCHMMA[.A-Z0-9_]*\|UTMALDG[.A-Z0-9_]*\|UTCBAR[.A-Z0-9_]*\|UTCATOMSWS[A-Z._0-9]*' | sort | uniq -c | sort -rn
} > profiles/sass_dense_gemm_$TAG.txt
{
  echo "# cuobjdump -sass $LIB — bm25_score_kernel<1536, true>"
  echo "# instruction counts:"
  kernel bm25_score_kernelILi1536ELb1 | grep -o 'FFMA\.RM\|LDG\.E\.NA\.64\.CONSTANT\|LDS\(\.128\)\?\|STS\(\.128\)\?\|MATCH\.ANY\|ATOMS[.A-Z]*\|ATOMG[.A-Z0-9]*\|RED[.A-Z0-9]*\|VOTE[.A-Z]*\|SHFL\.[A-Z]*\|BSSY\|STL\|LDL' | sort | uniq -c | sort -rn
  echo
  echo "# continuation rounds of a posting slice (apply_rest): four predicated 8-byte posting loads, then per posting"
  echo "# LOP3 (doc bits) + IMAD (accumulator address) + LDS + FFMA.RM + STS, class -> float (LEA.HI, FADD), FFMA.RM + ISETP (hit"
  echo "# test against the running bound), SEL / VIADD (hit registers) — no branch inside a round"
  L=$(kernel bm25_score_kernelILi1536ELb1 | grep -n 'FFMA.RM' | sed -n 9p | cut -d: -f1)
  kernel bm25_score_kernelILi1536ELb1 | sed -n "$((L-46)),$((L+14))p" | cut -c1-118
} > profiles/sass_bm25_score_$TAG.txt
{
  echo "# cuobjdump -sass $LIB — bm25_score16_kernel (two-phase BM25 score kernel, bm25_u16.cuh)"
  echo "# instruction counts:"
  kernel bm25_score16_kernel | grep -o 'FFMA\.RP\|FFMA\.RM\|LDG\.E\.NA\.64\.CONSTANT\|LDG\.E\.64\.CONSTANT\|CCTL[.A-Z0-9]*\|LDS\(\.U16\|\.128\)\?\|STS\(\.U16\|\.128\)\?\|MATCH\.ANY\|ATOMG[.A-Z0-9]*\|RED[.A-Z0-9]*\|VOTE[.A-Z]*\|SHFL\.[A-Z]*\|BSSY\|STL\|LDL' | sort | uniq -c | sort -rn
  echo
  echo "# phase 1, four interleaved rounds of one term (apply4): per posting LOP3 (doc bits) + IMAD (accumulator address) +"
  echo "# LDS.U16, FFMA.RP onto 2^23 + 1 (rounded-up 16-bit contribution), IADD3 (accumulate), STS.U16, SHF + IMAD (class penalty),"
  echo "# ISETP (hit test against the 16-bit bound), SEL / VIADD (the two hit registers) — no branch inside a round"
  L=$(kernel bm25_score16_kernel | grep -n 'FFMA.RP' | sed -n 2p | cut -d: -f1)
  kernel bm25_score16_kernel | sed -n "$((L-12)),$((L+66))p" | cut -c1-118
  echo
  echo "# phase 2 (flush): two binary searches in flight per lane, round-down FMA of the found impacts"
  L=$(kernel bm25_score16_kernel | grep -n 'LDG.E.64.CONSTANT' | sed -n 1p | cut -d: -f1)
  kernel bm25_score16_kernel | sed -n "$((L-14)),$((L+40))p" | cut -c1-118
} > profiles/sass_bm25_score16_$TAG.txt
rm -f $TMP
wc -l profiles/sass_dense_gemm_$TAG.txt profiles/sass_bm25_score_$TAG.txt
