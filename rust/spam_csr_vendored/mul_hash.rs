// Vendored spam_csr/src/mul_hash.rs — ONLY the body changes; the signature, the `Mul` impl
// (spam_csr/src/lib.rs:292-297) and the Matrix trait stay as they are.  The dispatch has to live
// inside spam_csr because CsrMatrix's fields are private (lib.rs:25-32).
//
// Not compiled in the build container (no Rust toolchain); this is the binding a maintainer adds.
use spam_cuda::DeviceScalar;
use spam_matrix::Matrix;

use crate::CsrMatrix;

impl<T: DeviceScalar + num_traits::NumAssign + Send + Sync, const B: bool> CsrMatrix<T, B> {
    // requires: rhs column indices be less than u32::MAX (unchanged, mul_hash.rs:12)
    pub fn mul_hash<const B1: bool, const B2: bool>(&self, rhs: &CsrMatrix<T, B1>) -> CsrMatrix<T, B2> {
        // B2 = true: rows sorted by column.  B2 = false: sorted rows are valid too (invariant6 only asks for distinct
        // columns, lib.rs:69-77) and cost one pass less; with the crate feature `reference-order` the device
        // permutes every row into the order the reference's own drain produces (mul_hash.rs:176-186).
        let (indices, vals, offsets) = spam_cuda::spgemm(
            self.rows.get(), self.cols.get(), &self.offsets, &self.indices, &self.vals,
            rhs.rows.get(), rhs.cols.get(), &rhs.offsets, &rhs.indices, &rhs.vals,
            !B2 && cfg!(feature = "reference-order"),
        );
        CsrMatrix { rows: self.rows, cols: rhs.cols, indices, vals, offsets }
    }
}
