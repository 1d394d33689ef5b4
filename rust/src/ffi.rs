//! 1:1 declaration of include/numrs_b200.h: the host-slice entry points, the device-buffer helpers and the
//! device-resident plan API.  The slab / IPC entry points are for the one-process-per-GPU
//! launchers (numrs_b200/dist_rlft3.py) and are not bound here: a Rust caller gets multi-GPU through "num_devices".
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_double, c_int};

pub const NRB_OK: c_int = 0;
pub const NRB_ERR_EMPTY_INPUT: c_int = -1;
pub const NRB_ERR_RESPONSE_TOO_LONG: c_int = -2;
pub const NRB_ERR_INVALID_ISIGN: c_int = -3;
pub const NRB_ERR_LENGTH_MISMATCH: c_int = -4;
pub const NRB_ERR_INVALID_DIMS: c_int = -5;
pub const NRB_ERR_ZERO_STDDEV: c_int = -8;
pub const NRB_PAD_LITERAL: c_int = 0;
pub const NRB_PAD_NR: c_int = 1;
pub const NRB_KIND_FOUR1: c_int = 1;
pub const NRB_KIND_FOURN: c_int = 2;
pub const NRB_KIND_REALFT: c_int = 3;
pub const NRB_KIND_RLFT3: c_int = 4;
pub const NRB_KIND_CONVLV: c_int = 5;
pub const NRB_KIND_CORREL: c_int = 6;

/// opaque plan handle (`nrb_plan_t`)
#[repr(C)]
pub struct nrb_plan_s {
    _private: [u8; 0],
}
pub type nrb_plan_t = *mut nrb_plan_s;

extern "C" {
    pub fn nrb_last_error() -> *const c_char;
    /// planner / runtime options, e.g. ("num_devices", 0): spread one host-slice call over every visible GPU
    pub fn nrb_set_option(name: *const c_char, value: std::os::raw::c_long) -> c_int;
    pub fn nrb_num_devices_in_use() -> c_int;
    pub fn nrb_shutdown() -> c_int;
    pub fn nrb_version() -> *const c_char;
    pub fn nrb_device_count() -> c_int;
    pub fn nrb_set_device(device: c_int) -> c_int;
    /// pinned host memory: host-slice calls and nrb_upload / nrb_download copy at full PCIe rate from / to it
    pub fn nrb_host_alloc(bytes: usize) -> *mut std::os::raw::c_void;
    pub fn nrb_host_free(p: *mut std::os::raw::c_void);
    /// device-resident plan API: `kind` = NRB_KIND_*, device pointers, caller's stream (null = default stream)
    pub fn nrb_plan_create(kind: c_int, dims: *const usize, ndim: usize, batch: usize, plan: *mut nrb_plan_t) -> c_int;
    pub fn nrb_plan_workspace_bytes(plan: nrb_plan_t) -> usize;
    pub fn nrb_plan_exec(plan: nrb_plan_t, d_io: *mut c_double, d_aux: *mut c_double, d_out: *mut c_double, isign: c_int,
                         arg: c_int, stream: *mut std::os::raw::c_void) -> c_int;
    pub fn nrb_plan_destroy(plan: nrb_plan_t) -> c_int;
    pub fn nrb_four1(data: *mut c_double, nn: usize, isign: c_int) -> c_int;
    pub fn nrb_four1_batch(ptrs: *const *mut c_double, nn: *const usize, count: usize, isign: c_int) -> c_int;
    pub fn nrb_fourn(data: *mut c_double, nn: *const usize, ndim: usize, isign: c_int) -> c_int;
    pub fn nrb_realft(data: *mut c_double, n: usize, isign: c_int) -> c_int;
    pub fn nrb_realft_batch(ptrs: *const *mut c_double, n: usize, count: usize, isign: c_int) -> c_int;
    pub fn nrb_rlft3(data: *mut c_double, speq: *mut c_double, nn1: usize, nn2: usize, nn3: usize, isign: c_int) -> c_int;
    pub fn nrb_convlv(data: *const c_double, n: usize, respns: *const c_double, m: usize, isign: c_int,
                      pad_mode: c_int, ans: *mut c_double) -> c_int;
    pub fn nrb_convlv_batch(data: *const *const c_double, count: usize, n: usize, respns: *const c_double, m: usize,
                            isign: c_int, pad_mode: c_int, ans: *const *mut c_double) -> c_int;
    pub fn nrb_correl(d1: *const c_double, n1: usize, d2: *const c_double, n2: usize, ans: *mut c_double) -> c_int;
    pub fn nrb_correl_batch(d1: *const *const c_double, d2: *const *const c_double, count: usize, n: usize,
                            ans: *const *mut c_double) -> c_int;
    pub fn nrb_correl_normalized(d1: *const c_double, n1: usize, d2: *const c_double, n2: usize, fast: c_int,
                                 ans: *mut c_double) -> c_int;
    pub fn nrb_autocorrel_fast(data: *const c_double, n: usize, ans: *mut c_double) -> c_int;
    pub fn nrb_twofft(d1: *const c_double, d2: *const c_double, n: usize, fft1: *mut c_double, fft2: *mut c_double) -> c_int;
    pub fn nrb_twofft_batch(d1: *const *const c_double, d2: *const *const c_double, count: usize, n: usize,
                            fft1: *const *mut c_double, fft2: *const *mut c_double) -> c_int;
    pub fn nrb_cosft1(y: *mut c_double, n: usize) -> c_int;
    pub fn nrb_cosft2(y: *mut c_double, n: usize, isign: c_int) -> c_int;
    pub fn nrb_sinft(y: *mut c_double, n: usize) -> c_int;
    pub fn nrb_device_alloc(bytes: usize, dptr: *mut *mut std::os::raw::c_void) -> c_int;
    pub fn nrb_device_free(dptr: *mut std::os::raw::c_void) -> c_int;
    pub fn nrb_upload(d_dst: *mut std::os::raw::c_void, h_src: *const std::os::raw::c_void, bytes: usize,
                      stream: *mut std::os::raw::c_void) -> c_int;
    pub fn nrb_download(h_dst: *mut std::os::raw::c_void, d_src: *const std::os::raw::c_void, bytes: usize,
                        stream: *mut std::os::raw::c_void) -> c_int;
    pub fn nrb_stream_synchronize(stream: *mut std::os::raw::c_void) -> c_int;
    pub fn nrb_complex_multiply_device(d_a: *mut c_double, d_b: *const c_double, ncomplex: usize, conj_b: c_int,
                                       scale: c_double, stream: *mut std::os::raw::c_void) -> c_int;
    pub fn nrb_power_spectrum(c: *const c_double, npoints: usize, take_sqrt: c_int, out: *mut c_double) -> c_int;
}

pub fn last_error() -> String {
    unsafe { std::ffi::CStr::from_ptr(nrb_last_error()).to_string_lossy().into_owned() }
}

/// four1 / realft / rlft3 return `()` in the reference and signal misuse by panicking.
pub fn panic_on(rc: c_int) {
    if rc != NRB_OK {
        panic!("numrs_b200: {}", last_error());
    }
}
