P=/root/repo/jpeg-xl-lossy-image-compression-thesis_b200
for v in A B; do JXLB200_LIB=$P/libjxlb200_$v.so python tools/exp_search.py; done
python tools/exp_search.py
