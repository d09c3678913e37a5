// NativeDnbc.scala -- JVM surface of libdnbc_b200.so (one instance per dnbc_ctx / per GPU).
//
// A maintainer of lucivpav/dnbc-scala adds this file to core/src/main/scala/ and replaces the bodies
// of DynamicNaiveBayesianClassifier.mle / inferMostLikelyHiddenStates / score with calls to it, as
// INTEGRATION.md spells out; signatures, model layout (the four learned maps, :14-17) and error
// behaviour (plain java.lang.Exception with the reference's messages) stay as they are.
// Natives are implemented in jni/dnbc_b200_jni.c over include/dnbc_b200.h.
// NOT compiled in this repository: the image has no JDK / scalac (SURVEY 8c).

final case class DenseParams(pi: Array[Double], a: Array[Double], b: Array[Double], w: Array[Double],
                             mu: Array[Double], variance: Array[Double], active: Array[Byte])

class NativeDnbc private (private var handle: Long, val nStates: Int, val alphabet: Array[Int], val mixture: Array[Int])
    extends AutoCloseable {
  import NativeDnbc._

  @native private def nDestroy(h: Long): Unit
  @native private def nLastError(h: Long): String
  @native private def nSetParams(h: Long, pi: Array[Double], a: Array[Double], b: Array[Double], w: Array[Double],
                                 mu: Array[Double], variance: Array[Double], active: Array[Byte]): Int
  @native private def nGetParams(h: Long, pi: Array[Double], a: Array[Double], b: Array[Double], w: Array[Double],
                                 mu: Array[Double], variance: Array[Double], active: Array[Byte]): Int
  @native private def nUpload(h: Long, nSeq: Long, offsets: Array[Long], disc: Array[Byte], cont: Array[Double],
                              labels: Array[Int]): Int
  @native private def nMle(h: Long, flags: Int, initCentroids: Array[Double], maxIters: Int, relTol: Double): Int
  @native private def nDecode(h: Long, mode: Int, precision: Int, path: Array[Short], score: Array[Double],
                              tie: Array[Byte]): Int
  @native private def nEmIterate(h: Long, iters: Int, loglik: Array[Double]): Int
  @native private def nStatsSize(h: Long): Long
  @native private def nEmEstep(h: Long, clampLabels: Boolean): Int
  @native private def nEmGetStats(h: Long, stats: Array[Double]): Int
  @native private def nEmSetStats(h: Long, stats: Array[Double]): Int
  @native private def nEmMstep(h: Long, totalSequences: Long): Int
  @native private def nCommInitRank(h: Long, nRanks: Int, rank: Int, id: Array[Byte]): Int
  @native private def nCommDestroy(h: Long): Int
  @native private def nEmAllreduce(h: Long): Int
  @native private def nEmIterateSharded(h: Long, iters: Int, totalSequences: Long, loglik: Array[Double]): Int

  private val vmax = if (alphabet.isEmpty) 1 else alphabet.max
  private val kmax = if (mixture.isEmpty) 1 else mixture.max
  private var points = 0
  private var sequences = 0

  // the reference signals every failure as `throw new Exception(message)` (DNBC.scala:46, 148, 274, 276)
  private def check(rc: Int): Unit = if (rc != 0) throw new Exception(nLastError(handle))

  def setParams(p: DenseParams): Unit = check(nSetParams(handle, p.pi, p.a, p.b, p.w, p.mu, p.variance, p.active))

  def getParams(): DenseParams = {
    val n = nStates; val d = alphabet.length; val c = mixture.length
    val p = DenseParams(new Array(n), new Array(n * n), new Array(d * n * vmax), new Array(c * n * kmax),
                        new Array(c * n * kmax), new Array(c * n * kmax), new Array(n))
    check(nGetParams(handle, p.pi, p.a, p.b, p.w, p.mu, p.variance, p.active))
    p
  }

  /** offsets i64 [S+1]; disc u8 [D][P] (id 255 = never-seen symbol); cont f64 [C][P]; labels i32 [P] or null */
  def upload(offsets: Array[Long], disc: Array[Byte], cont: Array[Double], labels: Array[Int]): Unit = {
    check(nUpload(handle, offsets.length - 1, offsets, disc, cont, labels))
    sequences = offsets.length - 1
    points = offsets.last.toInt
  }

  /** supervised mle on the resident, labelled sequences (the body of DynamicNaiveBayesianClassifier.mle) */
  def mle(flags: Int, initCentroids: Array[Double] = null): Unit = check(nMle(handle, flags, initCentroids, 0, 0.0))

  /** (path as state indices, per-sequence score, tie flags) for every resident sequence */
  def decode(mode: Int = DecodeReference, precision: Int = PrecF64): (Array[Short], Array[Double], Array[Byte]) = {
    val path = new Array[Short](points); val score = new Array[Double](sequences); val tie = new Array[Byte](points)
    check(nDecode(handle, mode, precision, path, score, tie))
    (path, score, tie)
  }

  /** Baum-Welch (additive API: the reference has no unsupervised learner); returns the log-likelihood series */
  def emIterate(iters: Int): Array[Double] = {
    val ll = new Array[Double](iters)
    check(nEmIterate(handle, iters, ll))
    ll
  }

  // multi-GPU, one executor process per GPU: the driver creates the id once (NativeDnbc.commUniqueId()),
  // ships it in a Spark broadcast variable, every executor joins and iterates on its own shard;
  // the statistics all-reduce is the library's (ncclAllReduce over NVLink, include/dnbc_b200.h)
  def commInitRank(nRanks: Int, rank: Int, id: Array[Byte]): Unit = check(nCommInitRank(handle, nRanks, rank, id))
  def commDestroy(): Unit = check(nCommDestroy(handle))
  def allreduce(): Unit = check(nEmAllreduce(handle))
  def emIterateSharded(iters: Int, totalSequences: Long): Array[Double] = {
    val ll = new Array[Double](iters)
    check(nEmIterateSharded(handle, iters, totalSequences, ll))
    ll
  }

  // finer-grained pieces of one iteration
  def statsSize: Int = nStatsSize(handle).toInt
  def estep(clampLabels: Boolean = false): Unit = check(nEmEstep(handle, clampLabels))
  def getStats(): Array[Double] = { val s = new Array[Double](statsSize); check(nEmGetStats(handle, s)); s }
  def setStats(s: Array[Double]): Unit = check(nEmSetStats(handle, s))
  def mstep(totalSequences: Long): Unit = check(nEmMstep(handle, totalSequences))

  override def close(): Unit = if (handle != 0) { nDestroy(handle); handle = 0 }
}

object NativeDnbc {
  System.loadLibrary("dnbc_b200_jni")

  val DecodeReference = 0; val DecodeBacktrace = 1          // DNBC_DECODE_*
  val PrecF32 = 0; val PrecF64 = 1                          // DNBC_PREC_*
  val QuirkTransitions = 1; val DoubleCountFirst = 2        // supervised-learner switches (SURVEY F3, F4)

  @native private def nCreate(nStates: Int, alphabet: Array[Int], mixture: Array[Int], device: Int): Long
  @native private def nCommUniqueId(id: Array[Byte]): Int
  @native private def nCommInitAll(handles: Array[Long]): Int
  @native private def nEmIterateGroup(handles: Array[Long], iters: Int, loglik: Array[Double]): Int

  /** 128-byte NCCL unique id for commInitRank (created on the driver, broadcast to the executors) */
  def commUniqueId(): Array[Byte] = {
    val id = new Array[Byte](128)
    if (nCommUniqueId(id) != 0) throw new Exception("dnbc_comm_unique_id failed")
    id
  }

  /** one JVM driving several GPUs (the reference's single driver thread): join one instance per device ... */
  def commInitAll(models: Seq[NativeDnbc]): Unit =
    if (nCommInitAll(models.map(_.handle).toArray) != 0) throw new Exception(models.head.nLastError(models.head.handle))

  /** ... and run Baum-Welch over all of them: local E-steps, one all-reduce, replicated M-steps per iteration */
  def emIterateGroup(models: Seq[NativeDnbc], iters: Int): Array[Double] = {
    val ll = new Array[Double](iters)
    if (nEmIterateGroup(models.map(_.handle).toArray, iters, ll) != 0)
      throw new Exception(models.head.nLastError(models.head.handle))
    ll
  }

  /** throws Exception("Number of suggested normal distributions ... at least one") for a hint < 1, like :276 */
  def create(nStates: Int, alphabet: Seq[Int], mixture: Seq[Int], device: Int = 0): NativeDnbc = {
    val al = alphabet.toArray; val mx = mixture.toArray
    new NativeDnbc(nCreate(nStates, al, mx, device), nStates, al, mx)
  }
}
