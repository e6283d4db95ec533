#!/bin/bash
# Launch lists of one training step (final build): configs[1] and the configs[2] shape, three eager steps each
# under ncu (gpu__time_duration, cold L2, serialised -> SHARES), summarised for the LAST step.
mkdir -p gpurun_out
L=gpurun_out/steps.log
echo "build $(cut -c1-12 lora_b200/.liblora_b200.stamp)" > $L
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file /tmp/c2_launches.csv python bench.py --profile-steps 3 > /dev/null 2>&1
echo "=== configs[1] step" >> $L
python scripts/summarize_launches.py /tmp/c2_launches.csv last >> $L 2>&1
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file /tmp/c3_launches.csv python bench.py --extended --rank 8 --profile-steps 3 > /dev/null 2>&1
echo "=== configs[2] shape step" >> $L
python scripts/summarize_launches.py /tmp/c3_launches.csv last >> $L 2>&1
cat $L
