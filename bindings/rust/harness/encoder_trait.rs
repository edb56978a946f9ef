//! `trait Encoder` for benchmark-jpegxl: lets `JXLCompressionBenchmark::run`
//! (benchmark-jpegxl/src/benchmark.rs:588-713) choose between the Docker `cjxl` it calls today
//! (`docker_manager.execute_cjxl`, benchmark.rs:654-660 -> docker_manager.rs:100-137) and the B200 library.
//!
//! Drop this file into `benchmark-jpegxl/src/encoder.rs`, add `pub mod encoder;` to `lib.rs`, and apply the three
//! edits listed at the bottom.  Source only: there is no Rust toolchain in the build image; the C ABI underneath is what
//! the test-suite exercises.

use std::error::Error;

use crate::docker_manager::DockerManager;

/// What one encode reports beyond the file it wrote.  `mp_per_s` and `bpp` become two new columns of `comparisons.csv`.
#[derive(Debug, Clone, Default)]
pub struct EncodeReport {
    pub stdout: String,
    pub codestream_bytes: u64,
    pub bpp: f64,       // 8 * bytes / (width * height); the harness derives it today as 24 / raw_file_size_ratio (benchmark.rs:921)
    pub encode_ms: f64, // device time of the encode (jxlb200_stats.total_ms) or wall time of the cjxl process
    pub mp_per_s: f64,  // width * height / 1e6 / (encode_ms / 1e3)
}

/// Same contract as `DockerManager::execute_cjxl` (docker_manager.rs:100-106): `Ok(Ok(report))` on success,
/// `Ok(Err(stderr))` when the encoder refuses the input — the caller's "skip" branch (benchmark.rs:661-677) is unchanged.
pub trait Encoder: Send {
    fn name(&self) -> &'static str;
    fn encode_file(&self, input_file: String, output_file: String, distance: f64, effort: u32)
        -> Result<Result<EncodeReport, String>, Box<dyn Error>>;
    /// true when `encode_file` already wrote `output_file` on the host (no `retrieve_file`, benchmark.rs:680-684)
    fn writes_host_file(&self) -> bool;
}

/// Today's path: libjxl's cjxl inside the worker's container.
pub struct DockerCjxl { pub docker: DockerManager }

impl Encoder for DockerCjxl {
    fn name(&self) -> &'static str { "docker-cjxl" }
    fn encode_file(&self, input_file: String, output_file: String, distance: f64, effort: u32)
        -> Result<Result<EncodeReport, String>, Box<dyn Error>> {
        let t0 = std::time::Instant::now();
        let r = self.docker.execute_cjxl(input_file, output_file, distance, effort)?;
        let ms = t0.elapsed().as_secs_f64() * 1e3;
        Ok(r.map(|stdout| EncodeReport { stdout, encode_ms: ms, ..Default::default() }))
    }
    fn writes_host_file(&self) -> bool { false }
}

/// The B200 path: `jxlb200::B200Encoder` (bindings/rust/jxlb200) behind the same trait.  `proposal` replaces
/// "which proposals/*.diff was applied before libjxl was rebuilt" (benchmark.rs:460-484): main -> 0,
/// homogeneity-partitioning.diff -> 1, homogeneity-factored-entropy.diff -> 2, combined.diff -> 3.
pub struct B200 { pub enc: jxlb200::B200Encoder }

impl Encoder for B200 {
    fn name(&self) -> &'static str { "b200" }
    fn encode_file(&self, input_file: String, output_file: String, distance: f64, effort: u32)
        -> Result<Result<EncodeReport, String>, Box<dyn Error>> {
        match self.enc.execute_cjxl_with_stats(input_file, output_file, distance, effort)? {
            Err(stderr) => Ok(Err(stderr)),
            Ok((stdout, s)) => {
                let mp = s.width as f64 * s.height as f64 / 1e6;
                Ok(Ok(EncodeReport { stdout, codestream_bytes: s.codestream_bytes, bpp: s.bpp, encode_ms: s.total_ms as f64,
                                     mp_per_s: mp / (s.total_ms as f64 / 1e3) }))
            }
        }
    }
    fn writes_host_file(&self) -> bool { true }
}

// ---------------------------------------------------------------------------------------------------------------
// Edits in benchmark-jpegxl (line numbers of the reference tree):
//
// 1. benchmark.rs:14-16   trait Benchmark { fn run(docker_manager: DockerManager, payload: &WorkerPayload) }
//      ->                 fn run(docker_manager: DockerManager, encoder: &dyn Encoder, payload: &WorkerPayload)
//    benchmark.rs:97-103  BenchmarkWorker::run builds the worker's encoder once: `B200::new(device = worker id % GPUs,
//                         proposal)` when `--encoder b200` (new clap flag next to main.rs:16-29), else `DockerCjxl`.
//                         With the B200 encoder the per-image clean / checkout / rebuild steps (benchmark.rs:452-484)
//                         are skipped: the proposal is a runtime enum.
//
// 2. benchmark.rs:654-660 `docker_manager.execute_cjxl(file, out_name, distance, effort)`
//      ->                 `encoder.encode_file(file, out_path_on_host_or_in_container, distance as f64, effort)`
//    benchmark.rs:680-684 `retrieve_file(...)` only `if !encoder.writes_host_file()`.
//
// 3. csv_writer.rs:24-42  ComparisonResult gains `pub bpp: f64, pub encode_ms: f64, pub mp_per_s: f64` (filled from the
//                         EncodeReport at benchmark.rs:702-710; the Docker path leaves bpp = 24 / raw_file_size_ratio);
//    csv_writer.rs:125-143 the header gains "bpp", "Encode ms", "MP/s" after "SSIMULACRA2" (appending keeps the first 17
//                         columns, so comparison_diffs.csv / summary.csv, csv_writer.rs:193-211, read old files unchanged);
//    csv_writer.rs:44-63  ComparisonResultDiff gains the matching `diff_bpp`, `diff_encode_ms`, `diff_mp_per_s`.
