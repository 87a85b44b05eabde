run base --trees 4096
clk base
export CUMIND_TEAM_TILE=1
run t4s7nw2 --trees 4096 --teams 4 --trees-per-team 7 --net-warps 2
unset CUMIND_TEAM_TILE
