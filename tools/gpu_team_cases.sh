clk t1_2048 --trees 2048
clk t2
run t3s10 --trees 4096 --teams 3 --trees-per-team 10
clk t3s10 --trees 4096 --teams 3 --trees-per-team 10
run t4s7 --trees 4096 --teams 4 --trees-per-team 7
run t2_nw5 --trees 4096 --net-warps 5
run t2_nw6 --trees 4096 --net-warps 6
run t2_nw8 --trees 4096 --net-warps 8
