for d in 0 1 2 3 4 7; do
  JPGENC_K3_DEBUG=$d ncu --metrics gpu__time_duration.sum --clock-control none -k regex:huffman_pack -c 3 --csv python tools/profile_run.py --iters 3 2>/dev/null | grep huffman | tail -1 | awk -F'","' -v d=$d '{print "debug="d, $NF}'
done
