timeout 600 python tools/dither_stress.py 2>&1 | tail -1
timeout 900 python tools/encode_clip.py --width 1920 --height 1080 --frames 600 --seq 75 --tiles 65536 --decode 1 2>&1 | tail -1 | cut -c1-1500
