for v in V1 V2; do echo "== $v"; ACRO_B200_LIB=$PWD/gymnast_optimalcontrol_b200/libacro_b200_$v.so timeout 200 python profiles/spec_probe.py 4096 50 spec1 2>&1 | grep "gamma_0=0.1"; done
