# usage: pipe_sweep.sh "nproc ctx waitUs blocking gates" ...   (diagnostic, GPU box)
W=/tmp/pp; python baseline/make_data.py c0 $W --reads 2000 >/dev/null; cd $W; /root/repo/baseline/_ref/sawritermc genome.sa genome.fa > /dev/null 2>&1
B=/root/repo/baseline/_ref
$B/blasrmc reads.fa genome.fa -sa genome.sa -sam -nproc 16 -out s.sam >/dev/null
for cfg in "$@"; do set -- $cfg
 s=$(date +%s.%N); BGPU_GATES=$5 BGPU_SERVICE_STATS=1 BGPU_SERVICE_CONTEXTS=$2 BGPU_BATCH_WAIT_US=$3 BGPU_BLOCKING_SYNC=$4 $B/blasrmc_gpu reads.fa genome.fa -sa genome.sa -sam -nproc $1 -out g.sam 2>&1 >/dev/null | grep RefineService; e=$(date +%s.%N); echo "nproc $1 ctx $2 wait $3 blocking $4 gates $5: $(python3 -c "print($e-$s)") s"
done
