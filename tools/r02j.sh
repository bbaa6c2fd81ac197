# compute-sanitizer (one tool per run of the smallest cases that exercise the hand-rolled shared-memory sorts,
# the pair rounds incl. their degenerate pairs, the bucket reduction and the NTT tile kernels)
set -x
T="python -m pytest -m gpu -x -q -p no:cacheprovider"
CASES="tests/test_gpu_msm_rounds.py::test_skewed tests/test_gpu_msm_rounds.py::test_kats_and_edge_cases tests/test_gpu_msm_rounds.py::test_two_stream_split_matches_one_batch tests/test_gpu_dft.py -k"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest -m gpu -x -q -p no:cacheprovider \
  "tests/test_gpu_msm_rounds.py::test_skewed" "tests/test_gpu_msm_rounds.py::test_kats_and_edge_cases" \
  "tests/test_gpu_msm_rounds.py::test_tables_and_ranges" "tests/test_gpu_kzg.py::test_commit_open_vs_oracle" \
  "tests/test_gpu_kzg.py::test_evaluations_on_a_domain_smaller_than_the_polynomial" \
  > gpurun_out/r02j_sanitizer_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r02j_sanitizer_memcheck.log
tail -5 gpurun_out/r02j_sanitizer_memcheck.log
