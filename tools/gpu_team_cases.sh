run t2 --schedule team --teams 2 --trees-per-team 16
clk t2 --schedule team --teams 2 --trees-per-team 16
run t4s8m2 --schedule team --teams 4 --trees-per-team 8 --net-warps 2
CUMIND_TEAM_TILE=1 run t4s8m2_tile42 --schedule team --teams 4 --trees-per-team 8 --net-warps 2
CUMIND_TEAM_TILE=1 clk t4s8m2_tile42 --schedule team --teams 4 --trees-per-team 8 --net-warps 2
CUMIND_TEAM_TILE=1 CUMIND_TEAM_STAGGER=2000 run t4s8m2_tile42_st --schedule team --teams 4 --trees-per-team 8 --net-warps 2
CUMIND_TEAM_TILE=1 run t3s8m2_tile42 --schedule team --teams 3 --trees-per-team 10 --net-warps 2
run big_team --schedule team --trees 32768
