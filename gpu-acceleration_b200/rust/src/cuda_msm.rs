//! `mopro_msm::msm::cuda_msm` -- Rust shim over libb200msm.so (include/b200msm.h).
//!
//! SOURCE ONLY: this image has no cargo/rustc, so this file is reviewed, not compiled here; the
//! same ABI is exercised by the C++ mirror (../cpp/cuda_msm.hpp) and the Python binding
//! (../b200msm.py), which the test-suite drives on the GPU.
//!
//! Drop-in for
//!     pub fn metal_variable_base_msm(bases: &[G1Affine], scalars: &[Fr])
//!         -> Result<G1Projective, Box<dyn Error>>
//!     (mopro-msm/src/msm/metal_msm/metal_msm.rs:642-695, re-exported at metal_msm/mod.rs:7)
//! Add to mopro-msm/src/msm/mod.rs:   #[cfg(feature = "cuda")] pub mod cuda_msm;
//! and re-export like the metal one:   pub use cuda_msm::cuda_variable_base_msm;
use ark_bn254::{Fq, Fr, G1Affine, G1Projective};
use ark_ff::BigInt;
use std::error::Error;
use std::ffi::CStr;
use std::mem::{offset_of, size_of};
use std::os::raw::{c_char, c_int, c_void};
use std::sync::OnceLock;

#[repr(C)]
pub struct B200MsmCtx {
    _private: [u8; 0],
}

#[link(name = "b200msm")]
extern "C" {
    fn b200msm_create(out: *mut *mut B200MsmCtx, devices: *const c_int, n_devices: c_int) -> c_int;
    fn b200msm_last_error(ctx: *const B200MsmCtx) -> *const c_char;
    fn b200msm_bn254_g1_msm(
        ctx: *mut B200MsmCtx,
        bases: *const c_void, base_stride: usize, x_off: usize, y_off: usize, inf_off: usize,
        scalars: *const c_void, scalar_stride: usize,
        n: usize, out_jacobian: *mut u64,
    ) -> c_int;
    fn b200msm_register_bases_ex(
        ctx: *mut B200MsmCtx,
        bases: *const c_void, base_stride: usize, x_off: usize, y_off: usize, inf_off: usize,
        n: usize, dev_indices: *const c_int, n_dev: c_int, precompute: c_int, out: *mut *mut B200MsmBases,
    ) -> c_int;
    fn b200msm_release_bases(ctx: *mut B200MsmCtx, h: *mut B200MsmBases) -> c_int;
    fn b200msm_bases_len(h: *const B200MsmBases) -> usize;
    fn b200msm_msm_registered(
        ctx: *mut B200MsmCtx, h: *const B200MsmBases,
        scalars: *const c_void, scalar_stride: usize, n: usize, out_jacobian: *mut u64,
    ) -> c_int;
}

extern "C" {
    fn b200msm_bn254_g2_msm(
        ctx: *mut B200MsmCtx,
        bases: *const c_void, base_stride: usize, x_off: usize, y_off: usize, inf_off: usize,
        scalars: *const c_void, scalar_stride: usize,
        n: usize, out_jacobian: *mut u64,
    ) -> c_int;
}

#[repr(C)]
pub struct B200MsmBases {
    _private: [u8; 0],
}

struct Ctx(*mut B200MsmCtx);
// The C context serialises internally (one MSM at a time per context).
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {}

/// Process-global context, created on first use (the reference rebuilds its pipeline on every call,
/// metal_msm.rs:693; this is the persistent replacement).  It spans the CUDA devices listed in the environment variable
/// `B200MSM_DEVICES` (comma-separated ordinals, e.g. "0,1,2,3": every MSM is then sharded by point range over them);
/// without the variable it is device 0 only -- a library must not grab every GPU of a shared host by default.
fn default_ctx() -> Result<&'static Ctx, Box<dyn Error>> {
    static CTX: OnceLock<Result<Ctx, String>> = OnceLock::new();
    CTX.get_or_init(|| unsafe {
        let mut p: *mut B200MsmCtx = std::ptr::null_mut();
        let devices: Vec<c_int> = std::env::var("B200MSM_DEVICES")
            .map(|v| v.split(',').filter_map(|t| t.trim().parse().ok()).collect())
            .unwrap_or_default();
        let rc = b200msm_create(&mut p, if devices.is_empty() { std::ptr::null() } else { devices.as_ptr() }, devices.len() as c_int);
        if rc != 0 {
            Err(CStr::from_ptr(b200msm_last_error(std::ptr::null())).to_string_lossy().into_owned())
        } else {
            Ok(Ctx(p))
        }
    })
    .as_ref()
    .map_err(|e| e.clone().into())
}

/// Same contract as `metal_variable_base_msm`: empty input is an error, unequal lengths are
/// truncated to the shorter, the result equals `G1Projective::msm(bases, scalars)` as a group element.
pub fn cuda_variable_base_msm(
    mut bases: &[G1Affine],
    mut scalars: &[Fr],
) -> Result<G1Projective, Box<dyn Error>> {
    if bases.is_empty() || scalars.is_empty() {
        return Err("Empty input".into()); // metal_msm.rs:647-649
    }
    if bases.len() != scalars.len() {
        let n = std::cmp::min(bases.len(), scalars.len()); // metal_msm.rs:652-656
        bases = &bases[..n];
        scalars = &scalars[..n];
    }
    let ctx = default_ctx()?;
    // `G1Affine` is not repr(C): its layout is MEASURED here and passed through the ABI, so the
    // device repack kernel reads x, y and the infinity flag wherever rustc put them.
    // Fq / Fr are `Fp<MontBackend<_, 4>, 4>(BigInt<4>([u64; 4]))`: 32 bytes, Montgomery form, LE limbs.
    let mut out = [0u64; 12];
    let rc = unsafe {
        b200msm_bn254_g1_msm(
            ctx.0,
            bases.as_ptr() as *const c_void,
            size_of::<G1Affine>(),
            offset_of!(G1Affine, x),
            offset_of!(G1Affine, y),
            offset_of!(G1Affine, infinity),
            scalars.as_ptr() as *const c_void,
            size_of::<Fr>(),
            bases.len(),
            out.as_mut_ptr(),
        )
    };
    if rc != 0 {
        let msg = unsafe { CStr::from_ptr(b200msm_last_error(ctx.0)) }.to_string_lossy().into_owned();
        return Err(msg.into());
    }
    // The library returns fully reduced Montgomery words: build the field elements without conversion.
    let fq = |w: &[u64]| Fq::new_unchecked(BigInt::new([w[0], w[1], w[2], w[3]]));
    Ok(G1Projective::new_unchecked(fq(&out[0..4]), fq(&out[4..8]), fq(&out[8..12])))
}

/// The B2 multi-scalar multiplication of a Groth16 prover (the reference has no G2 path).  Same contract as the G1 call.
/// `Fq2` is `QuadExtField { c0, c1 }`: the library reads x at `offset_of!(G2Affine, x)` as c0 || c1 (32 bytes each).
pub fn cuda_variable_base_msm_g2(
    mut bases: &[ark_bn254::G2Affine],
    mut scalars: &[Fr],
) -> Result<ark_bn254::G2Projective, Box<dyn Error>> {
    use ark_bn254::{Fq2, G2Affine, G2Projective};
    if bases.is_empty() || scalars.is_empty() {
        return Err("Empty input".into());
    }
    let n = std::cmp::min(bases.len(), scalars.len());
    bases = &bases[..n];
    scalars = &scalars[..n];
    let ctx = default_ctx()?;
    let mut out = [0u64; 24];
    let rc = unsafe {
        b200msm_bn254_g2_msm(
            ctx.0,
            bases.as_ptr() as *const c_void,
            size_of::<G2Affine>(),
            offset_of!(G2Affine, x),
            offset_of!(G2Affine, y),
            offset_of!(G2Affine, infinity),
            scalars.as_ptr() as *const c_void,
            size_of::<Fr>(),
            n,
            out.as_mut_ptr(),
        )
    };
    if rc != 0 {
        return Err(last_error(ctx));
    }
    let fq = |w: &[u64]| Fq::new_unchecked(BigInt::new([w[0], w[1], w[2], w[3]]));
    let fq2 = |w: &[u64]| Fq2::new(fq(&w[0..4]), fq(&w[4..8]));
    Ok(G2Projective::new_unchecked(fq2(&out[0..8]), fq2(&out[8..16]), fq2(&out[16..24])))
}

fn last_error(ctx: &Ctx) -> Box<dyn Error> {
    unsafe { CStr::from_ptr(b200msm_last_error(ctx.0)) }.to_string_lossy().into_owned().into()
}

/// A base set kept on the GPU(s) across MSMs -- the proving-key pattern (a Groth16 prover runs the A, B1, C, H MSMs
/// of every proof over the same points).  `precompute = true` also builds the one-time window table
/// `2^(c*w) * P_i` (W x 64 bytes of HBM per point) so that each MSM needs one bucket reduce and no Horner step.
/// Nothing like it exists in the reference, which re-uploads and re-converts the points on every call
/// (metal_msm.rs:74-201).
pub struct RegisteredBases {
    ctx: &'static Ctx,
    handle: *mut B200MsmBases,
}
unsafe impl Send for RegisteredBases {}

impl RegisteredBases {
    pub fn new(bases: &[G1Affine], precompute: bool) -> Result<Self, Box<dyn Error>> {
        if bases.is_empty() {
            return Err("Empty input".into());
        }
        let ctx = default_ctx()?;
        let mut handle: *mut B200MsmBases = std::ptr::null_mut();
        let rc = unsafe {
            b200msm_register_bases_ex(
                ctx.0,
                bases.as_ptr() as *const c_void,
                size_of::<G1Affine>(),
                offset_of!(G1Affine, x),
                offset_of!(G1Affine, y),
                offset_of!(G1Affine, infinity),
                bases.len(),
                std::ptr::null(),
                0,
                precompute as c_int,
                &mut handle,
            )
        };
        if rc != 0 {
            return Err(last_error(ctx));
        }
        Ok(Self { ctx, handle })
    }

    pub fn len(&self) -> usize {
        unsafe { b200msm_bases_len(self.handle) }
    }

    /// `sum_i scalars[i] * bases[i]` over the first `min(scalars.len(), self.len())` registered points.
    pub fn msm(&self, scalars: &[Fr]) -> Result<G1Projective, Box<dyn Error>> {
        if scalars.is_empty() {
            return Err("Empty input".into());
        }
        let n = std::cmp::min(scalars.len(), self.len());
        let mut out = [0u64; 12];
        let rc = unsafe {
            b200msm_msm_registered(self.ctx.0, self.handle, scalars.as_ptr() as *const c_void, size_of::<Fr>(), n, out.as_mut_ptr())
        };
        if rc != 0 {
            return Err(last_error(self.ctx));
        }
        let fq = |w: &[u64]| Fq::new_unchecked(BigInt::new([w[0], w[1], w[2], w[3]]));
        Ok(G1Projective::new_unchecked(fq(&out[0..4]), fq(&out[4..8]), fq(&out[8..12])))
    }
}

impl Drop for RegisteredBases {
    fn drop(&mut self) {
        unsafe { b200msm_release_bases(self.ctx.0, self.handle) };
    }
}

#[cfg(test)]
mod tests {
    use super::*;
    use ark_ec::VariableBaseMSM;
    // Mirrors tests/cuzk/e2e.rs:14-63 with the CUDA entry point.
    #[test]
    fn test_e2e_cuda_msm_pipeline() {
        let (bases, scalars) = crate::msm::metal_msm::test_utils::generate_random_bases_and_scalars(1 << 16);
        let got = cuda_variable_base_msm(&bases, &scalars).unwrap();
        let want = G1Projective::msm(&bases, &scalars).unwrap();
        assert_eq!(got, want);
    }

    #[test]
    fn test_registered_bases_with_table() {
        let (bases, scalars) = crate::msm::metal_msm::test_utils::generate_random_bases_and_scalars(1 << 16);
        let want = G1Projective::msm(&bases, &scalars).unwrap();
        for precompute in [false, true] {
            let key = RegisteredBases::new(&bases, precompute).unwrap();
            assert_eq!(key.msm(&scalars).unwrap(), want);
        }
    }
}
