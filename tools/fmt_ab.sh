sed -E 's/C=1 std_table=False: //; s/repeat identical //; s/cfg5 stack ms \(median of 5\)/ms/'
