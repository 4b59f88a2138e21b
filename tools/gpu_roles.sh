mkdir -p gpurun_out
{
for spec in "8 1024 1024 32 32 fwd" "8 512 512 64 64 fwd" "8 512 512 64 32 fup" "8 256 256 128 64 fup" "8 256 256 128 128 fwd"; do
  SFK_FLAGS=518 python tests/prof_igemm.py $spec 5
  SFK_FLAGS=518 SFK_ROLES=1 python tests/prof_igemm.py $spec 5
done
for spec in "8 1024 1024 32 32 dgrad" "8 512 512 64 64 dgrad" "8 512 512 32 64 fupb" "8 256 256 128 128 dgrad"; do
  SFK_FLAGS=0 python tests/prof_igemm.py $spec 5
  SFK_FLAGS=0 SFK_ROLES=1 python tests/prof_igemm.py $spec 5
done
python tests/prof_elem.py 8 1024 32
python tests/prof_elem.py 8 512 64
python tests/prof_elem.py 8 256 128
} > gpurun_out/r2_roles.log 2>&1
echo done
