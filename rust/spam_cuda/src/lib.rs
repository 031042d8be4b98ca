//! Thin FFI over libspam_cuda.so (include/spam_cuda.h).  One `Handle` per thread (it owns a CUDA
//! stream and workspace and is not thread-safe); a non-zero status becomes a panic in the safe
//! wrappers, mirroring the reference's own panics (spam_csr/src/mul_hash.rs:47-48,190; lib.rs:270).
#![allow(non_camel_case_types)]
use std::{cell::RefCell, ffi::CStr, os::raw::{c_char, c_int, c_void}};

#[repr(C)] pub struct spam_handle { _p: [u8; 0] }

pub const SPAM_F32: c_int = 0;
pub const SPAM_F64: c_int = 1;
pub const SPAM_I32: c_int = 2;
pub const SPAM_I64: c_int = 3;

extern "C" {
    pub fn spam_cuda_create(h: *mut *mut spam_handle, device: c_int) -> c_int;
    pub fn spam_cuda_destroy(h: *mut spam_handle) -> c_int;
    pub fn spam_strerror(status: c_int) -> *const c_char;
    pub fn spam_last_error(h: *const spam_handle) -> *const c_char;
    pub fn spam_spgemm_symbolic(h: *mut spam_handle, dtype: c_int, a_rows: u64, a_cols: u64, a_ptr: *const u64,
        a_idx: *const u64, a_val: *const c_void, b_rows: u64, b_cols: u64, b_ptr: *const u64, b_idx: *const u64,
        b_val: *const c_void, c_ptr: *mut u64, c_nnz: *mut u64) -> c_int;
    /// sorted = 1: rows sorted by column (B2 = true, and a valid B2 = false result); sorted = 0: the reference's own
    /// B2 = false order (slot order of linprobe's map, mul_hash.rs:176-186)
    pub fn spam_spgemm_numeric(h: *mut spam_handle, c_idx: *mut u64, c_val: *mut c_void, sorted: c_int) -> c_int;
    /// page-locked host memory for callers that want full-rate PCIe copies (ordinary Vecs are staged by the library)
    pub fn spam_host_alloc(p: *mut *mut c_void, bytes: u64) -> c_int;
    pub fn spam_host_free(p: *mut c_void) -> c_int;
    /// several GPUs of one node, one process per GPU (collective calls; see include/spam_cuda.h)
    pub fn spam_comm_unique_id(out128: *mut c_void) -> c_int;
    pub fn spam_comm_init(h: *mut spam_handle, id128: *const c_void, rank: c_int, world: c_int) -> c_int;
    pub fn spam_comm_destroy(h: *mut spam_handle) -> c_int;
    pub fn spam_spmv(h: *mut spam_handle, dtype: c_int, a_rows: u64, a_cols: u64, a_ptr: *const u64, a_idx: *const u64,
        a_val: *const c_void, x: *const c_void, y: *mut c_void) -> c_int;
    pub fn spam_dok_to_csr(h: *mut spam_handle, dtype: c_int, rows: u64, cols: u64, n: u64, tri_rows: *const u64,
        tri_cols: *const u64, tri_vals: *const c_void, c_ptr: *mut u64, c_nnz: *mut u64) -> c_int;
    pub fn spam_dok_to_csr_fetch(h: *mut spam_handle, c_idx: *mut u64, c_val: *mut c_void) -> c_int;
    /// `Matrix::transpose` of `CsrMatrix` (spam_csr/src/lib.rs:256-264); t_ptr: cols+1, t_idx/t_val: nnz entries
    pub fn spam_csr_transpose(h: *mut spam_handle, dtype: c_int, rows: u64, cols: u64, ptr: *const u64, idx: *const u64,
        val: *const c_void, t_ptr: *mut u64, t_idx: *mut u64, t_val: *mut c_void) -> c_int;
    /// `impl Add / Sub for CsrMatrix` -> apply_elementwise (spam_csr/src/lib.rs:83-149): op 0 add, 1 sub, `| 2` for
    /// IS_SORTED = false operands; phase 1 writes c_ptr[rows+1] and *c_nnz, spam_csr_ewise_fetch fills idx / val
    pub fn spam_csr_ewise(h: *mut spam_handle, op: c_int, dtype: c_int, rows: u64, cols: u64, a_ptr: *const u64,
        a_idx: *const u64, a_val: *const c_void, b_ptr: *const u64, b_idx: *const u64, b_val: *const c_void,
        c_ptr: *mut u64, c_nnz: *mut u64) -> c_int;
    pub fn spam_csr_ewise_fetch(h: *mut spam_handle, c_idx: *mut u64, c_val: *mut c_void) -> c_int;
    /// parse_matrix_market (spam_dok/src/lib.rs:282-478) as a triplet stream for spam_dok_to_csr
    pub fn spam_mm_parse(text: *const c_char, len: u64, out: *mut spam_mm) -> c_int;
    pub fn spam_mm_free(m: *mut spam_mm);
}

/// Mirror of `struct spam_mm` (include/spam_cuda.h).
#[repr(C)]
pub struct spam_mm {
    pub kind: c_int,
    pub rows: u64,
    pub cols: u64,
    pub declared_entries: u64,
    pub n: u64,
    pub tri_rows: *mut u64,
    pub tri_cols: *mut u64,
    pub tri_vals: *mut c_void,
    pub err: [c_char; 96],
}

/// Element types the device serves.  `Wrapping<i32/i64>` are `repr(transparent)` and map to I32/I64.
/// Sealed: there is no CPU fallback, so other `T` do not get a `mul_hash`.
pub trait DeviceScalar: Copy + sealed::Sealed { const DTYPE: c_int; }
mod sealed { pub trait Sealed {} }
macro_rules! scalar { ($t:ty, $d:expr) => { impl sealed::Sealed for $t {} impl DeviceScalar for $t { const DTYPE: c_int = $d; } } }
scalar!(f32, SPAM_F32); scalar!(f64, SPAM_F64); scalar!(i32, SPAM_I32); scalar!(i64, SPAM_I64);
scalar!(std::num::Wrapping<i32>, SPAM_I32); scalar!(std::num::Wrapping<i64>, SPAM_I64);

pub struct Handle(*mut spam_handle);
impl Handle {
    pub fn new(device: i32) -> Self {
        let mut h = std::ptr::null_mut();
        let st = unsafe { spam_cuda_create(&mut h, device) };
        assert!(st == 0, "spam_cuda_create: {}", unsafe { CStr::from_ptr(spam_strerror(st)) }.to_string_lossy());
        Handle(h)
    }
    fn check(&self, st: c_int) {
        if st != 0 {
            let (a, b) = unsafe { (CStr::from_ptr(spam_strerror(st)), CStr::from_ptr(spam_last_error(self.0))) };
            panic!("spam_cuda: {} ({})", a.to_string_lossy(), b.to_string_lossy());
        }
    }
}
impl Drop for Handle { fn drop(&mut self) { unsafe { spam_cuda_destroy(self.0); } } }
thread_local! { static HANDLE: RefCell<Option<Handle>> = RefCell::new(None); }

/// C = A * B on the GPU.  Slices are the CsrMatrix fields (offsets, indices, vals); returns the same
/// three vectors for C, cancellation zeros kept; rows sorted by column, or (`reference_order`) in the order the
/// reference's B2 = false branch emits them.  The Vecs are ordinary pageable memory: the library moves them through
/// its own ring of pinned slots (hostio.cu).
pub fn spgemm<T: DeviceScalar>(a_rows: usize, a_cols: usize, a_off: &[usize], a_idx: &[usize], a_val: &[T],
                               b_rows: usize, b_cols: usize, b_off: &[usize], b_idx: &[usize], b_val: &[T],
                               reference_order: bool) -> (Vec<usize>, Vec<T>, Vec<usize>) {
    const _: () = assert!(std::mem::size_of::<usize>() == 8);
    HANDLE.with(|cell| {
        let mut slot = cell.borrow_mut();
        let h = slot.get_or_insert_with(|| Handle::new(0));
        let mut offsets: Vec<usize> = vec![0; a_rows + 1];
        let mut nnz = 0u64;
        h.check(unsafe { spam_spgemm_symbolic(h.0, T::DTYPE, a_rows as u64, a_cols as u64, a_off.as_ptr() as *const u64,
            a_idx.as_ptr() as *const u64, a_val.as_ptr() as *const c_void, b_rows as u64, b_cols as u64,
            b_off.as_ptr() as *const u64, b_idx.as_ptr() as *const u64, b_val.as_ptr() as *const c_void,
            offsets.as_mut_ptr() as *mut u64, &mut nnz) });
        // Vec::with_capacity(nnz) + set_len, exactly like mul_hash.rs:119,196-199
        let (mut indices, mut vals): (Vec<usize>, Vec<T>) = (Vec::with_capacity(nnz as usize), Vec::with_capacity(nnz as usize));
        // reference_order: B2 = false with the reference's slot order; otherwise rows sorted by column (valid for both B2)
        h.check(unsafe { spam_spgemm_numeric(h.0, indices.as_mut_ptr() as *mut u64, vals.as_mut_ptr() as *mut c_void,
                                             if reference_order { 0 } else { 1 }) });
        unsafe { indices.set_len(nnz as usize); vals.set_len(nnz as usize); }
        (indices, vals, offsets)
    })
}
