for b in 512 1024 2048 3000 4096 4736 6144 8192; do run auto_$b --trees $b; run warp_$b --schedule warp --trees $b; done
run atari_auto --workload atari --trees 1024
run atari_warp --workload atari --trees 1024 --schedule warp
