cp h264-scroll-encoder_b200/libh264scroll_b200.so /tmp/orig.so
cp variants/tr1.so h264-scroll-encoder_b200/libh264scroll_b200.so
python bench.py --steps 1 --warmup 0 --no-cpu --no-extras --no-check 2>&1 | tail -2 | cut -c1-300
ls -la gpurun_out/trace.bin
cp /tmp/orig.so h264-scroll-encoder_b200/libh264scroll_b200.so
