timeout 300 python -m pytest tests/test_gpu_train.py tests/test_gpu_loss_head.py tests/test_gpu_configs.py -x -q 2>&1 | tail -3
one() { timeout 200 python bench.py --steps 40 --no-cpu-baseline --no-fp32-mode --no-profile 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1 %.4f ms/step e2e %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step']))"; }
one seg3; VCM_EARLY_UPDATE=0 one no_early; one seg3
