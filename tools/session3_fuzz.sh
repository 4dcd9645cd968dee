# fuzz battery on the final library (session 3): GPU vs oracle, vs the reference's renderer, every render path vs every other, textured meshes
O=gpurun_out/s3_fuzz.txt
: > $O
( timeout 200 python tools/fuzz_poses.py 150 31 2>&1 | tail -1 | sed 's/^/fuzz_poses.py 150 31: /' ) >> $O
( timeout 200 python tools/fuzz_reference.py 150 32 2>&1 | tail -1 | sed 's/^/fuzz_reference.py 150 32: /' ) >> $O
( timeout 200 python tools/fuzz_paths.py 100 33 2>&1 | tail -1 | sed 's/^/fuzz_paths.py 100 33: /' ) >> $O
( timeout 200 python tools/fuzz_mesh.py 60 34 2>&1 | tail -1 | sed 's/^/fuzz_mesh.py 60 34: /' ) >> $O
( timeout 200 python tools/fuzz_state.py 2>&1 | tail -1 | sed 's/^/fuzz_state.py: /' ) >> $O
cat $O
