#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_next_rows_gpu.py tests/test_audio_model_gpu.py tests/test_image_embedding_gpu.py tests/test_tc_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -80 ) > gpurun_out/pytest_round3.log 2>&1
( time timeout 200 python scripts/wn_train_probe.py ) > gpurun_out/wn_train.log 2>&1
tail -4 gpurun_out/pytest_round3.log; grep "^{" gpurun_out/wn_train.log | cut -c1-250
