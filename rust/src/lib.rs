//! num_rs -- drop-in for the FFT hot path of SciRustaceans/numrs, running on a B200 through
//! libnumrs_b200 (include/numrs_b200.h).  Public signatures are the reference's, unchanged:
//! interleaved f64 complex, nn/ndim/isign conventions, unnormalised inverse.
#![allow(non_snake_case)]

pub mod ffi;

/// Spread every host-slice call (`rlft3`, 3-D `Fourn`, `fft_batch`, `convlv_batch`, `correl_batch`) over `n` GPUs of the
/// box inside this process (0 = every visible device, 1 = the calling thread's device: the default).  Not part of the
/// reference's API: the one knob a deployment sets once; every reference signature stays as it is.
pub fn set_num_devices(n: usize) {
    let rc = unsafe { ffi::nrb_set_option(b"num_devices\0".as_ptr() as *const std::os::raw::c_char, n as std::os::raw::c_long) };
    assert!(rc == ffi::NRB_OK, "nrb_set_option(num_devices) failed");
}

pub mod FFT_1 {
    use crate::ffi::*;

    /// reference: src/FFT_1.rs:5
    pub fn four1(data: &mut [f64], nn: usize, isign: i32) {
        assert!(data.len() >= 2 * nn, "index out of bounds");
        panic_on(unsafe { nrb_four1(data.as_mut_ptr(), nn, isign) });
    }

    /// reference: src/FFT_1.rs:110 (numerically equivalent variant)
    pub fn four1_optimized(data: &mut [f64], nn: usize, isign: i32) {
        four1(data, nn, isign)
    }

    /// reference: src/FFT_1.rs:143-190; builder flags accepted, one GPU path
    pub struct FFTProcessor {
        max_threads: usize,
        use_optimized: bool,
    }

    impl FFTProcessor {
        pub fn new() -> Self { Self { max_threads: 1, use_optimized: true } }
        pub fn with_threads(mut self, threads: usize) -> Self { self.max_threads = threads; self }
        pub fn with_optimized(mut self, use_optimized: bool) -> Self { self.use_optimized = use_optimized; self }
        pub fn fft(&self, data: &mut [f64], isign: i32) {
            let nn = data.len() / 2;
            four1(data, nn, isign)
        }
        /// whole batch in one device call (pointer / length tables across the ABI)
        pub fn fft_batch(&self, batches: &mut [&mut [f64]], isign: i32) {
            let ptrs: Vec<*mut f64> = batches.iter_mut().map(|b| b.as_mut_ptr()).collect();
            let nn: Vec<usize> = batches.iter().map(|b| b.len() / 2).collect();
            panic_on(unsafe { nrb_four1_batch(ptrs.as_ptr(), nn.as_ptr(), ptrs.len(), isign) });
        }
    }

    pub fn real_to_complex(real_data: &[f64]) -> Vec<f64> {
        real_data.iter().flat_map(|&v| [v, 0.0]).collect()
    }
    pub fn complex_to_real(complex_data: &[f64]) -> Vec<f64> {
        complex_data.iter().step_by(2).copied().collect()
    }
    pub fn power_spectrum(c: &[f64]) -> Vec<f64> {
        c.chunks(2).map(|p| if p.len() == 2 { p[0] * p[0] + p[1] * p[1] } else { 0.0 }).collect()
    }
    pub fn magnitude_spectrum(c: &[f64]) -> Vec<f64> {
        power_spectrum(c).into_iter().map(f64::sqrt).collect()
    }
}

pub mod Fourn {
    use crate::ffi::*;
    use std::io::{Error, ErrorKind};

    /// The in-memory N-dimensional FFT the reference calls at src/Real_FT3.rs:35 but never
    /// defines; validation as src/Fourn.rs:367-378.
    #[allow(non_snake_case)]
    pub fn Fourn(data: &mut [f64], nn: &[usize], ndim: usize, isign: i32) {
        fourn(data, nn, ndim, isign).expect("Fourn")
    }

    pub fn fourn(data: &mut [f64], nn: &[usize], ndim: usize, isign: i32) -> Result<(), Error> {
        if ndim == 0 || ndim > nn.len() {
            return Err(Error::new(ErrorKind::InvalidInput, "Invalid dimensions"));
        }
        let total = nn[..ndim]
            .iter()
            .try_fold(1usize, |acc, &d| acc.checked_mul(d))
            .ok_or_else(|| Error::new(ErrorKind::InvalidInput, "Invalid dimensions"))?;
        if total == 0 {
            return Err(Error::new(ErrorKind::InvalidInput, "Empty dimensions"));
        }
        // nrb_fourn copies exactly 2 * total doubles from and to the slice: a shorter slice must panic here (as the
        // reference's indexing would), never reach the C side
        let needed = total.checked_mul(2).expect("2 * prod(nn) overflows usize");
        assert!(data.len() >= needed, "data length must be at least 2 * prod(nn)");
        let rc = unsafe { nrb_fourn(data.as_mut_ptr(), nn.as_ptr(), ndim, isign) };
        match rc {
            NRB_OK => Ok(()),
            NRB_ERR_INVALID_DIMS | NRB_ERR_INVALID_ISIGN => Err(Error::new(ErrorKind::InvalidInput, last_error())),
            _ => Err(Error::new(ErrorKind::Other, last_error())),
        }
    }
}

pub mod Real_FT {
    use crate::ffi::*;

    /// reference: src/Real_FT.rs:4
    pub fn realft(data: &mut [f64], n: usize, isign: i32) {
        assert!(n % 2 == 0, "n must be even");
        assert!(data.len() >= n, "data length must be at least n");
        panic_on(unsafe { nrb_realft(data.as_mut_ptr(), n, isign) });
    }
    pub fn realft_optimized(data: &mut [f64], n: usize, isign: i32) { realft(data, n, isign) }

    /// reference: src/Real_FT.rs:332-370
    pub struct RealFTProcessor { use_optimized: bool, parallel_threshold: usize }
    impl RealFTProcessor {
        pub fn new() -> Self { Self { use_optimized: true, parallel_threshold: 1024 } }
        pub fn with_optimized(mut self, v: bool) -> Self { self.use_optimized = v; self }
        pub fn with_threshold(mut self, t: usize) -> Self { self.parallel_threshold = t; self }
        pub fn process(&self, data: &mut [f64], n: usize, isign: i32) { realft(data, n, isign) }
        pub fn process_batch(&self, batches: &mut [(&mut [f64], usize, i32)]) {
            // equal (n, isign) runs go to the device as one batch
            let mut i = 0;
            while i < batches.len() {
                let (n, isign) = (batches[i].1, if batches[i].2 == 1 { 1 } else { -1 });
                let mut j = i;
                let mut ptrs = Vec::new();
                while j < batches.len() && batches[j].1 == n && (if batches[j].2 == 1 { 1 } else { -1 }) == isign {
                    assert!(n % 2 == 0 && batches[j].0.len() >= n);
                    ptrs.push(batches[j].0.as_mut_ptr());
                    j += 1;
                }
                panic_on(unsafe { nrb_realft_batch(ptrs.as_ptr(), n, ptrs.len(), isign) });
                i = j;
            }
        }
    }
}

pub mod Real_FT3 {
    use crate::ffi::*;
    use ndarray::prelude::*;

    /// reference: src/Real_FT3.rs:8
    pub fn rlft3(data: &mut Array3<f64>, speq: &mut Array2<f64>, nn1: usize, nn2: usize, nn3: usize, isign: i32) {
        assert!(isign == 1 || isign == -1, "isign must be 1 or -1");
        assert!(data.shape() == &[nn1, nn2, nn3], "data dimensions mismatch");
        assert!(speq.shape() == &[nn1, 2 * nn2], "speq dimensions mismatch");
        let d = data.as_slice_mut().unwrap();
        let s = speq.as_slice_mut().unwrap();
        panic_on(unsafe { nrb_rlft3(d.as_mut_ptr(), s.as_mut_ptr(), nn1, nn2, nn3, isign) });
    }

    /// reference: src/Real_FT3.rs:145 (flat slices)
    pub fn rlft3_optimized(data: &mut [f64], speq: &mut [f64], nn1: usize, nn2: usize, nn3: usize, isign: i32) {
        assert!(isign == 1 || isign == -1);
        assert!(data.len() == nn1 * nn2 * nn3 && speq.len() == nn1 * 2 * nn2);
        panic_on(unsafe { nrb_rlft3(data.as_mut_ptr(), speq.as_mut_ptr(), nn1, nn2, nn3, isign) });
    }
}

pub mod Convolve {
    use crate::ffi::*;
    use ndarray::prelude::*;

    /// reference: src/Convolve.rs:226-238
    #[derive(Debug, thiserror::Error)]
    pub enum ConvlvError {
        #[error("Input arrays cannot be empty")]
        EmptyInput,
        #[error("Response function longer than data")]
        ResponseTooLong,
        #[error("isign must be 1 (convolution) or -1 (deconvolution)")]
        InvalidIsign,
        #[error("Division by zero in deconvolution")]
        DivisionByZero,
        #[error("FFT computation error: {0}")]
        FftError(String),
    }

    fn map(rc: i32) -> Result<(), ConvlvError> {
        match rc {
            NRB_OK => Ok(()),
            NRB_ERR_EMPTY_INPUT => Err(ConvlvError::EmptyInput),
            NRB_ERR_RESPONSE_TOO_LONG => Err(ConvlvError::ResponseTooLong),
            NRB_ERR_INVALID_ISIGN => Err(ConvlvError::InvalidIsign),
            _ => Err(ConvlvError::FftError(last_error())),
        }
    }

    /// reference: src/Convolve.rs:8
    pub fn convlv(data: &[f64], respns: &[f64], isign: i32) -> Result<Array1<f64>, ConvlvError> {
        let mut ans = vec![0.0f64; data.len()];
        map(unsafe {
            nrb_convlv(data.as_ptr(), data.len(), respns.as_ptr(), respns.len(), isign, NRB_PAD_LITERAL, ans.as_mut_ptr())
        })?;
        Ok(Array1::from_vec(ans))
    }

    /// reference: src/Convolve.rs:241; equal-length signals go to the device as one batch and
    /// the response spectrum is computed once
    pub fn convlv_batch(data_batch: &[&[f64]], respns: &[f64], isign: i32) -> Result<Vec<Array1<f64>>, ConvlvError> {
        if data_batch.is_empty() { return Ok(Vec::new()); }
        let n = data_batch[0].len();
        if data_batch.iter().any(|d| d.len() != n) {
            return data_batch.iter().map(|d| convlv(d, respns, isign)).collect();
        }
        let mut outs: Vec<Vec<f64>> = data_batch.iter().map(|_| vec![0.0; n]).collect();
        let ip: Vec<*const f64> = data_batch.iter().map(|d| d.as_ptr()).collect();
        let op: Vec<*mut f64> = outs.iter_mut().map(|o| o.as_mut_ptr()).collect();
        map(unsafe {
            nrb_convlv_batch(ip.as_ptr(), ip.len(), n, respns.as_ptr(), respns.len(), isign, NRB_PAD_LITERAL, op.as_ptr())
        })?;
        Ok(outs.into_iter().map(Array1::from_vec).collect())
    }

    /// reference: src/Convolve.rs:253-339
    pub struct ConvlvProcessor { use_optimized: bool, parallel_threshold: usize }
    impl ConvlvProcessor {
        pub fn new() -> Self { Self { use_optimized: true, parallel_threshold: 1024 } }
        pub fn with_optimized(mut self, v: bool) -> Self { self.use_optimized = v; self }
        pub fn with_threshold(mut self, t: usize) -> Self { self.parallel_threshold = t; self }
        pub fn process(&self, data: &[f64], respns: &[f64], isign: i32) -> Result<Array1<f64>, ConvlvError> {
            convlv(data, respns, isign)
        }
        pub fn process_batch(&self, b: &[&[f64]], respns: &[f64], isign: i32) -> Result<Vec<Array1<f64>>, ConvlvError> {
            convlv_batch(b, respns, isign)
        }
    }
}

pub mod Correlation {
    use crate::ffi::*;
    use ndarray::prelude::*;

    /// reference: src/Correlation.rs:389-399
    #[derive(Debug, thiserror::Error)]
    pub enum CorrelError {
        #[error("Input arrays cannot be empty")]
        EmptyInput,
        #[error("Input arrays must have the same length")]
        LengthMismatch,
        #[error("FFT computation error: {0}")]
        FftError(String),
        #[error("Normalization error: standard deviation is zero")]
        ZeroStdDev,
    }

    fn map(rc: i32) -> Result<(), CorrelError> {
        match rc {
            NRB_OK => Ok(()),
            NRB_ERR_EMPTY_INPUT => Err(CorrelError::EmptyInput),
            NRB_ERR_LENGTH_MISMATCH => Err(CorrelError::LengthMismatch),
            NRB_ERR_ZERO_STDDEV => Err(CorrelError::ZeroStdDev),
            _ => Err(CorrelError::FftError(last_error())),
        }
    }

    /// reference: src/Correlation.rs:8
    pub fn correl(data1: &[f64], data2: &[f64]) -> Result<Array1<f64>, CorrelError> {
        let mut ans = vec![0.0f64; data1.len()];
        map(unsafe { nrb_correl(data1.as_ptr(), data1.len(), data2.as_ptr(), data2.len(), ans.as_mut_ptr()) })?;
        Ok(Array1::from_vec(ans))
    }

    /// reference: src/Correlation.rs:273
    pub fn correl_batch(data_pairs: &[(&[f64], &[f64])]) -> Result<Vec<Array1<f64>>, CorrelError> {
        if data_pairs.is_empty() { return Ok(Vec::new()); }
        let n = data_pairs[0].0.len();
        if n == 0 || data_pairs.iter().any(|(a, b)| a.len() != n || b.len() != n) {
            return data_pairs.iter().map(|(a, b)| correl(a, b)).collect();
        }
        let mut outs: Vec<Vec<f64>> = data_pairs.iter().map(|_| vec![0.0; n]).collect();
        let ap: Vec<*const f64> = data_pairs.iter().map(|p| p.0.as_ptr()).collect();
        let bp: Vec<*const f64> = data_pairs.iter().map(|p| p.1.as_ptr()).collect();
        let op: Vec<*mut f64> = outs.iter_mut().map(|o| o.as_mut_ptr()).collect();
        map(unsafe { nrb_correl_batch(ap.as_ptr(), bp.as_ptr(), ap.len(), n, op.as_ptr()) })?;
        Ok(outs.into_iter().map(Array1::from_vec).collect())
    }

    /// reference: src/Correlation.rs:281
    pub fn autocorrel(data: &[f64]) -> Result<Array1<f64>, CorrelError> { correl(data, data) }

    fn normalized(data1: &[f64], data2: &[f64], fast: i32) -> Result<Array1<f64>, CorrelError> {
        let mut ans = vec![0.0f64; data1.len()];
        map(unsafe { nrb_correl_normalized(data1.as_ptr(), data1.len(), data2.as_ptr(), data2.len(), fast, ans.as_mut_ptr()) })?;
        Ok(Array1::from_vec(ans))
    }
    /// reference: src/Correlation.rs:189
    pub fn correl_normalized(data1: &[f64], data2: &[f64]) -> Result<Array1<f64>, CorrelError> { normalized(data1, data2, 0) }
    /// reference: src/Correlation.rs:226
    pub fn correl_normalized_fast(data1: &[f64], data2: &[f64]) -> Result<Array1<f64>, CorrelError> { normalized(data1, data2, 1) }
    /// reference: src/Correlation.rs:286
    pub fn autocorrel_fast(data: &[f64]) -> Result<Array1<f64>, CorrelError> {
        let mut ans = vec![0.0f64; data.len()];
        map(unsafe { nrb_autocorrel_fast(data.as_ptr(), data.len(), ans.as_mut_ptr()) })?;
        Ok(Array1::from_vec(ans))
    }
}

pub mod Cos_FT {
    use crate::ffi::*;
    /// reference: src/Cos_FT.rs:7 (1-based array: y[0] unused, data y[1..=n+1])
    pub fn cosft1(y: &mut [f64], n: usize) {
        assert!(y.len() >= n + 2, "index out of bounds: y must hold n + 2 elements");
        panic_on(unsafe { nrb_cosft1(y.as_mut_ptr(), n) });
    }
    /// reference: src/Cos_FT.rs:77
    pub fn cosft1_optimized(y: &mut [f64], n: usize) { cosft1(y, n) }
}

pub mod Cos_FT2 {
    use crate::ffi::*;
    /// reference: src/Cos_FT2.rs:7
    pub fn cosft2(y: &mut [f64], n: usize, isign: i32) {
        if isign != 1 && isign != -1 { panic!("Invalid isign value: {}. Must be 1 or -1", isign); }
        assert!(y.len() >= n + 1, "index out of bounds: y must hold n + 1 elements");
        panic_on(unsafe { nrb_cosft2(y.as_mut_ptr(), n, isign) });
    }
    /// reference: src/Cos_FT2.rs:202 (same contract)
    pub fn cosft2_simd(y: &mut [f64], n: usize, isign: i32) { cosft2(y, n, isign) }
}

pub mod Sin_FT {
    use crate::ffi::*;
    /// README.md:72 lists sinft; the reference has no source file for it (NR semantics)
    pub fn sinft(y: &mut [f64], n: usize) {
        assert!(y.len() >= n + 1, "index out of bounds: y must hold n + 1 elements");
        panic_on(unsafe { nrb_sinft(y.as_mut_ptr(), n) });
    }
}

/// Device-resident buffers (not in the reference: SURVEY.md 8f N1) so chains of transforms stay in HBM.
pub mod device {
    use crate::ffi::*;
    use std::os::raw::c_void;
    pub struct DeviceBuffer { ptr: *mut c_void, len: usize }
    impl DeviceBuffer {
        pub fn new(len: usize) -> Self {
            let mut ptr: *mut c_void = std::ptr::null_mut();
            panic_on(unsafe { nrb_device_alloc(8 * len.max(1), &mut ptr) });
            Self { ptr, len }
        }
        pub fn from_slice(host: &[f64]) -> Self { let b = Self::new(host.len()); b.upload(host); b }
        pub fn upload(&self, host: &[f64]) {
            assert!(host.len() <= self.len);
            panic_on(unsafe { nrb_upload(self.ptr, host.as_ptr() as *const c_void, 8 * host.len(), std::ptr::null_mut()) });
            panic_on(unsafe { nrb_stream_synchronize(std::ptr::null_mut()) });
        }
        pub fn download(&self) -> Vec<f64> {
            let mut host = vec![0.0f64; self.len];
            panic_on(unsafe { nrb_download(host.as_mut_ptr() as *mut c_void, self.ptr, 8 * self.len, std::ptr::null_mut()) });
            panic_on(unsafe { nrb_stream_synchronize(std::ptr::null_mut()) });
            host
        }
        pub fn as_mut_ptr(&self) -> *mut f64 { self.ptr as *mut f64 }
        pub fn len(&self) -> usize { self.len }
    }
    impl Drop for DeviceBuffer { fn drop(&mut self) { unsafe { nrb_device_free(self.ptr); } } }
}

pub mod FFT_2 {
    use crate::ffi::*;
    /// reference: src/FFT_2.rs:3 (asserts :5-7)
    pub fn twofft(data1: &[f64], data2: &[f64], fft1: &mut [f64], fft2: &mut [f64]) {
        let n = data1.len();
        assert_eq!(data2.len(), n, "data2 length must equal data1 length");
        assert_eq!(fft1.len(), 2 * n + 2, "fft1 must have length 2*n + 2");
        assert_eq!(fft2.len(), 2 * n + 2, "fft2 must have length 2*n + 2");
        panic_on(unsafe { nrb_twofft(data1.as_ptr(), data2.as_ptr(), n, fft1.as_mut_ptr(), fft2.as_mut_ptr()) });
    }
    /// reference: src/FFT_2.rs:135
    pub fn twofft_optimized(data1: &[f64], data2: &[f64], fft1: &mut [f64], fft2: &mut [f64]) { twofft(data1, data2, fft1, fft2) }

    /// reference: src/FFT_2.rs:222-265 (builder flags kept and ignored: one device path)
    pub struct TwoFFTProcessor { use_optimized: bool, parallel_threshold: usize }
    impl TwoFFTProcessor {
        pub fn new() -> Self { Self { use_optimized: true, parallel_threshold: 1024 } }
        pub fn with_optimized(mut self, v: bool) -> Self { self.use_optimized = v; self }
        pub fn with_threshold(mut self, t: usize) -> Self { self.parallel_threshold = t; self }
        pub fn process(&self, data1: &[f64], data2: &[f64], fft1: &mut [f64], fft2: &mut [f64]) { twofft(data1, data2, fft1, fft2) }
        /// reference: src/FFT_2.rs:258 (which cannot mutate through its `&[(.., &mut [f64], &mut [f64])]` argument and
        /// does not compile, err.log; the slice of tuples is taken mutably here).  Runs of equal length = one batch.
        pub fn process_batch(&self, batches: &mut [(&[f64], &[f64], &mut [f64], &mut [f64])]) {
            let mut i = 0;
            while i < batches.len() {
                let n = batches[i].0.len();
                let (mut a, mut b, mut f1, mut f2) = (Vec::new(), Vec::new(), Vec::new(), Vec::new());
                let mut j = i;
                while j < batches.len() && batches[j].0.len() == n {
                    assert_eq!(batches[j].1.len(), n, "data2 length must equal data1 length");
                    assert_eq!(batches[j].2.len(), 2 * n + 2, "fft1 must have length 2*n + 2");
                    assert_eq!(batches[j].3.len(), 2 * n + 2, "fft2 must have length 2*n + 2");
                    a.push(batches[j].0.as_ptr()); b.push(batches[j].1.as_ptr());
                    f1.push(batches[j].2.as_mut_ptr()); f2.push(batches[j].3.as_mut_ptr());
                    j += 1;
                }
                panic_on(unsafe { nrb_twofft_batch(a.as_ptr(), b.as_ptr(), a.len(), n, f1.as_ptr(), f2.as_ptr()) });
                i = j;
            }
        }
    }

    /// reference: src/FFT_2.rs:361
    pub fn extract_real_imag(fft: &[f64]) -> (Vec<f64>, Vec<f64>) {
        let n = fft.len() / 2;
        ((0..n).map(|i| fft[2 * i]).collect(), (0..n).map(|i| fft[2 * i + 1]).collect())
    }
    /// reference: src/FFT_2.rs:375
    pub fn combine_real_imag(real: &[f64], imag: &[f64]) -> Vec<f64> {
        assert_eq!(real.len(), imag.len(), "Real and imaginary parts must have same length");
        real.iter().zip(imag.iter()).flat_map(|(r, i)| [*r, *i]).collect()
    }
}
