set -x
O=gpurun_out
python tools/profile_loss.py --mode disp > $O/r3d_profile.txt 2>&1
STV_LIB=$PWD/slowtv_monodepth_b200/csrc/build/libstv_nopf.so python tools/profile_loss.py --mode disp > $O/r3d_profile_nopf.txt 2>&1
python tools/profile_loss.py --mode disp > $O/r3d_profile_b.txt 2>&1
STV_LIB=$PWD/slowtv_monodepth_b200/csrc/build/libstv_nopf.so python tools/profile_loss.py --mode disp > $O/r3d_profile_nopf_b.txt 2>&1
grep -E "photo_fwd" $O/r3d_profile.txt $O/r3d_profile_nopf.txt $O/r3d_profile_b.txt $O/r3d_profile_nopf_b.txt
