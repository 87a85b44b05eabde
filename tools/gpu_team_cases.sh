run t2 --schedule team --teams 2 --trees-per-team 16
clk t2 --schedule team --teams 2 --trees-per-team 16
