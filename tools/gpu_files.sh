#!/bin/bash
TAG=${1:-files}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python bench.py --stage files --config 2 --lnc 200 > $OUT/files_c2_200.json 2> $OUT/files_c2_200.err; echo "rc=$?"; cat $OUT/files_c2_200.json | head -c 1500; echo
timeout 600 python bench.py --stage files --config 2 --lnc 1000 > $OUT/files_c2_1000.json 2> $OUT/files_c2_1000.err; echo "rc=$?"; cat $OUT/files_c2_1000.json | head -c 1500; echo
timeout 900 python bench.py --stage files --config 5 --lnc 20000 > $OUT/files_c5_20000.json 2> $OUT/files_c5_20000.err; echo "rc=$?"; cat $OUT/files_c5_20000.json | head -c 1500; echo
timeout 300 python bench.py --stage ingest --lnc 96 > $OUT/ingest.json 2> $OUT/ingest.err; echo "rc=$?"; cat $OUT/ingest.json | head -c 1200; echo
