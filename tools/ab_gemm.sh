for shape in "1 36864 2304 768 NT b16" "512 384 384 384 NN b16" "512 384 384 384 NT f32" "512 2304 384 96 NT f32" "1 147456 768 2304 NN f32" "1 8192 8192 8192 NT b16"; do
  for v in v2 ""; do
      echo -n "variant=${v:-default} : "; MIRROR_B200_VARIANT=$v python tools/gemm_one.py $shape
  done
done
