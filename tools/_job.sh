for b in sort_check sort_check_t; do echo $b; for c in "1000000 0 1 0" "2000000 0 1 0" "4194304 0 1 0" "8000000 0 1 0"; do build/$b $c; done; done
