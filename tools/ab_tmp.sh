#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python tools/fill_timing.py > $O/r02y_fill_timing5.log 2>&1; echo "timing rc=$?"; cat $O/r02y_fill_timing5.log | cut -c1-400
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "eri or fill or h2_known or cart_to_sph" > $O/r02y_fill_pytest6.log 2>&1; echo "pytest rc=$? $(tail -n 1 $O/r02y_fill_pytest6.log)"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_fill_scatter -c 12 --csv --log-file $O/r02y_fill_scatter_ncu3.csv python tools/fill_timing.py n2_ccpvtz ne2_uhf_ccpvqz > $O/r02y_fill_ncu4.log 2>&1; echo "ncu rc=$?"
