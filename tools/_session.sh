D=gpurun_out/${1:-t1}; mkdir -p $D
timeout 1200 python tools/tune_gather.py --skip-copy --out $D/tune_xform.jsonl > $D/tune_xform.txt 2>&1; echo "tune rc=$?"
timeout 600 python tools/tune_gather.py --skip-copy --translate --patches 256,448 --out $D/tune_xform_tr.jsonl > $D/tune_xform_tr.txt 2>&1; echo "tune tr rc=$?"
timeout 600 python tools/tune_gather.py --skip-copy --crop-gb 0.2 --patches 256 --out $D/tune_xform_small.jsonl > $D/tune_xform_small.txt 2>&1; echo "tune small rc=$?"
timeout 600 python tools/tune_gather.py --skip-copy --crop-gb 0.2 --translate --patches 256 --out $D/tune_xform_small_tr.jsonl > $D/tune_xform_small_tr.txt 2>&1; echo "tune small tr rc=$?"
echo done
