for cfg in "TCN_NO_PDL=1 TCN_NO_WGRAD_STREAM=1" "A=1"; do
  echo "== $cfg"; env $cfg python tools/exp/pdl_graph_check.py 800 2 2>/dev/null | awk '{print $1, $NF}' | tr '\n' ';'; echo
done
