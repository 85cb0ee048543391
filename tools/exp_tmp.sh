for v in "" "-DNB_X_NO_SEQ" "-DNB_X_NO_P3" "-DNB_X_NO_P3 -DNB_X_NO_2B"; do
  HIMUT_B200_NVCC_EXTRA="$v" python -c "import __graft_entry__ as g; g.build(force=True, load=False)" > /dev/null 2>&1
  echo "variant [$v]"
  python tools/run_path.py normcounts --contig-mb 64 --reps 2 2>&1 | tail -1 | cut -c1-200
done
