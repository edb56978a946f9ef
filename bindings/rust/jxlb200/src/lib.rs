//! Drop-in for the encoder call of `benchmark-jpegxl`:
//! `DockerManager::execute_cjxl(input_file, output_file, distance, effort)` (docker_manager.rs:100-137) +
//! `retrieve_file` (docker_manager.rs:72-87) become one FFI call into `libjxlb200.so`.
use jxlb200_sys as sys;
use std::error::Error;
use std::ffi::CStr;

pub use sys::{
    JXLB200_FLAG_QUALITY, JXLB200_PROPOSAL_COMBINED, JXLB200_PROPOSAL_FACTORED_ENTROPY, JXLB200_PROPOSAL_NONE,
    JXLB200_PROPOSAL_PARTITIONING,
};

/// What the device already knows after an encode (the harness derives bpp as 24 / raw_file_size_ratio,
/// benchmark.rs:921; PSNR as calculate_psnr, image_reader.rs:604-606).
#[derive(Clone, Copy, Debug)]
/// What `execute_cjxl_with_stats` reports next to the text line.
#[derive(Debug, Clone)]
pub struct EncodeStatsFull { pub codestream_bytes: u64, pub bpp: f64, pub total_ms: f32, pub width: u32, pub height: u32 }

pub struct EncodeStats {
    pub codestream_bytes: u64,
    pub bpp: f64,
    pub device_ms: f32,
    pub psnr: Option<f64>,
    pub width: u32,
    pub height: u32,
}

/// One per worker thread, like each worker's `DockerManager` clone (benchmark.rs:97-103).  A context is owned by
/// one thread; distinct contexts run concurrently.
pub struct B200Encoder {
    ctx: *mut sys::jxlb200_ctx,
    proposal: u32,
    flags: u32,
}
unsafe impl Send for B200Encoder {}

impl B200Encoder {
    /// `proposal`: which `proposals/*.diff` the reference would have applied before rebuilding libjxl
    /// (benchmark.rs:460-484); "main" is `JXLB200_PROPOSAL_NONE`.
    pub fn new(device: i32, proposal: u32) -> Result<Self, String> {
        // one hardware queue per pipeline stream; must be set before the first CUDA call of the process
        if std::env::var_os("CUDA_DEVICE_MAX_CONNECTIONS").is_none() {
            std::env::set_var("CUDA_DEVICE_MAX_CONNECTIONS", "32");
        }
        if unsafe { sys::jxlb200_abi_version() } != sys::JXLB200_ABI_VERSION {
            return Err("libjxlb200.so has a different ABI version".into());
        }
        let ctx = unsafe { sys::jxlb200_create(device) };
        if ctx.is_null() {
            return Err("no sm_100 CUDA device (there is no CPU fallback)".into());
        }
        Ok(Self { ctx, proposal, flags: 0 })
    }

    /// Also fill `EncodeStats::psnr` from the device's own reconstruction of the coded frame.
    pub fn with_quality(mut self) -> Self {
        self.flags |= sys::JXLB200_FLAG_QUALITY;
        self
    }

    /// `cjxl`'s defaults have the Gaborish loop filter and chroma-from-luma fitting on; here they are opt-in
    /// (`JXLB200_FLAG_GABORISH`, `JXLB200_FLAG_CFL`: this library's own sharpening kernel and fit).
    pub fn with_loop_filter_and_cfl(mut self, gaborish: bool, cfl: bool) -> Self {
        if gaborish { self.flags |= sys::JXLB200_FLAG_GABORISH; }
        if cfl { self.flags |= sys::JXLB200_FLAG_CFL; }
        self
    }

    fn last_error(&self) -> String {
        unsafe { CStr::from_ptr(sys::jxlb200_last_error(self.ctx)) }.to_string_lossy().into_owned()
    }

    /// In-memory form: interleaved RGB8 in, codestream bytes out.
    pub fn encode_rgb8(&self, pixels: &[u8], width: u32, height: u32, distance: f32, effort: u32)
        -> Result<(Vec<u8>, EncodeStats), String> {
        if pixels.len() < 3 * width as usize * height as usize {
            return Err("pixel buffer smaller than 3*width*height".into());
        }
        let image = sys::jxlb200_image { pixels: pixels.as_ptr(), width, height, stride: 3 * width as usize };
        let params = sys::jxlb200_params { distance, effort, proposal: self.proposal, flags: self.flags };
        let (mut out, mut len) = (std::ptr::null_mut(), 0usize);
        let mut stats = std::mem::MaybeUninit::<sys::jxlb200_stats>::zeroed();
        let rc = unsafe { sys::jxlb200_encode(self.ctx, &image, &params, &mut out, &mut len, stats.as_mut_ptr()) };
        if rc != 0 {
            return Err(self.last_error());
        }
        let bytes = unsafe { std::slice::from_raw_parts(out, len) }.to_vec();
        unsafe { sys::jxlb200_free(out as *mut _) };
        let s = unsafe { stats.assume_init() };
        let psnr = if s.quality_valid != 0 { Some(s.psnr) } else { None };
        Ok((bytes, EncodeStats { codestream_bytes: s.codestream_bytes, bpp: s.bpp, device_ms: s.total_ms, psnr, width, height }))
    }

    /// `execute_cjxl` that also hands back the statistics (used by the harness' `trait Encoder`,
    /// bindings/rust/harness/encoder_trait.rs, to fill the bpp / MP/s columns of comparisons.csv).
    pub fn execute_cjxl_with_stats(&self, input_file: String, output_file: String, distance: f64, effort: u32)
        -> Result<Result<(String, EncodeStatsFull), String>, Box<dyn Error>> {
        let img = image::open(&input_file)?.to_rgb8();
        let (w, h) = img.dimensions();
        match self.encode_rgb8(img.as_raw(), w, h, distance as f32, effort) {
            Err(msg) => Ok(Err(msg)),
            Ok((bytes, s)) => {
                std::fs::write(&output_file, &bytes)?;
                let text = format!("Compressed to {} bytes ({:.3} bpp) in {:.3} ms", s.codestream_bytes, s.bpp, s.device_ms);
                Ok(Ok((text, EncodeStatsFull { codestream_bytes: s.codestream_bytes, bpp: s.bpp, total_ms: s.device_ms, width: w, height: h })))
            }
        }
    }

    /// Same contract as `DockerManager::execute_cjxl` (docker_manager.rs:100-106): `Ok(Ok(stdout))` on success,
    /// `Ok(Err(stderr))` when the encoder refuses the input, so the caller's "skip" branch
    /// (benchmark.rs:661-677) keeps working; `Err` only for I/O failures of the host side.
    pub fn execute_cjxl(&self, input_file: String, output_file: String, distance: f64, effort: u32)
        -> Result<Result<String, String>, Box<dyn Error>> {
        let img = image::open(&input_file)?.to_rgb8();
        let (w, h) = img.dimensions();
        match self.encode_rgb8(img.as_raw(), w, h, distance as f32, effort) {
            Err(msg) => Ok(Err(msg)),
            Ok((bytes, s)) => {
                std::fs::write(&output_file, &bytes)?; // replaces `docker cp` (docker_manager.rs:72-87)
                Ok(Ok(format!("Compressed to {} bytes ({:.3} bpp) in {:.3} ms", s.codestream_bytes, s.bpp, s.device_ms)))
            }
        }
    }
}

impl Drop for B200Encoder {
    fn drop(&mut self) {
        unsafe { sys::jxlb200_destroy(self.ctx) }
    }
}
