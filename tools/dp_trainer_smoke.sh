#!/bin/bash
# Data-parallel smoke of the SR trainer entry point (N = number of GPUs on the box, default 2):
#   bash tools/dp_trainer_smoke.sh [N]
n=${1:-2}
out=gpurun_out
mkdir -p $out
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29531 \
  Super_resolution/code/train_adaptive_unet.py --scale 0.5 --depth_override 2 --patch_size 32 --batch_size 8 --epochs 2 \
  --patches_per_image 4 --synthetic 8 --precision bf16 --model_dir /tmp/dp_models --log_dir /tmp/dp_logs --run_name dp \
  > $out/dp_trainer.log 2>&1
echo "dp trainer rc=$?"
grep -E "Epoch|loss:|Training complete|PSNR|Error|error" $out/dp_trainer.log | head -20
ls /tmp/dp_models /tmp/dp_logs/dp 2>/dev/null | head
