python tools/encode_clip.py --decode 0 2>&1 | tail -1 | cut -c1-420
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/encode_clip.py --sharded 1 2>&1 | tail -1 | cut -c1-700
