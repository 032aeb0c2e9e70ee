cd /root/repo
for i in 1 2 3; do
echo new; python tools/time_pair_hidden.py 48 256 | grep bwd
echo old; DFOL_LIB_PATH=/root/repo/tools/_old/libdfol_b200.so python tools/time_pair_hidden.py 48 256 | grep bwd
done
echo new; python tools/time_pair_hidden.py 64 256 | grep bwd
echo old; DFOL_LIB_PATH=/root/repo/tools/_old/libdfol_b200.so python tools/time_pair_hidden.py 64 256 | grep bwd
