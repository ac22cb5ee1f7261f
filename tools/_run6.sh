set -x
python -m pytest tests -m gpu -x -q > gpurun_out/a6_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/a6_pytest.log
tail -5 gpurun_out/a6_pytest.log
python bench.py --families-only > gpurun_out/a6_fam_sb4.txt 2>&1
SMOS_MSDA_SB=2 python bench.py --families-only > gpurun_out/a6_fam_sb2.txt 2>&1
SMOS_POOL_FOLD=0 python bench.py --families-only > gpurun_out/a6_fam_fold0.txt 2>&1
tail -1 gpurun_out/a6_fam_*.txt
SMOS_FAMILIES_RAW=1 python bench.py --families-only > gpurun_out/a6_fam_raw.txt 2>&1; tail -1 gpurun_out/a6_fam_raw.txt
