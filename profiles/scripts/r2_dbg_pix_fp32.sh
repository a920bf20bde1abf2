F='grep -v "^\*\|OMP_NUM"'
VARIANT=default timeout 100 python profiles/dbg_pix_fp32.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -5
VARIANT=eager FQL_B200_GRAPH=0 timeout 100 python profiles/dbg_pix_fp32.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -5
VARIANT=blocking FQL_B200_GRAPH=0 CUDA_LAUNCH_BLOCKING=1 timeout 100 python profiles/dbg_pix_fp32.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -5
VARIANT=memcheck FQL_B200_GRAPH=0 timeout 250 compute-sanitizer --tool memcheck --print-limit 5 python profiles/dbg_pix_fp32.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -30
