# usage: pipe_sweep.sh "nproc contexts threads" ...   (diagnostic, GPU box): wall time of blasrmc_gpu on configs[0] (2000 reads)
W=/tmp/pp; python baseline/make_data.py c0 $W --reads 2000 >/dev/null; cd $W; /root/repo/baseline/_ref/sawritermc genome.sa genome.fa > /dev/null 2>&1
B=/root/repo/baseline/_ref
s=$(date +%s.%N); $B/blasrmc reads.fa genome.fa -sa genome.sa -sam -nproc 16 -out s.sam >/dev/null; e=$(date +%s.%N); echo "stock nproc 16 (cold): $(python3 -c "print($e-$s)") s"
s=$(date +%s.%N); $B/blasrmc reads.fa genome.fa -sa genome.sa -sam -nproc 16 -out s.sam >/dev/null; e=$(date +%s.%N); echo "stock nproc 16: $(python3 -c "print($e-$s)") s"
grep -v "^@PG" s.sam | sort | md5sum
for cfg in "$@"; do set -- $cfg
 s=$(date +%s.%N); BGPU_SERVICE_STATS=1 BGPU_SERVICE_CONTEXTS=$2 BGPU_THREADS=$3 $B/blasrmc_gpu reads.fa genome.fa -sa genome.sa -sam -nproc $1 -out g.sam 2>&1 >/dev/null | grep -E "RefineService|BGPU_HOST" | tail -8; e=$(date +%s.%N); echo "fibers $1 ctx $2 threads $3: $(python3 -c "print($e-$s)") s  $(grep -v "^@PG" g.sam | sort | md5sum)"
done
