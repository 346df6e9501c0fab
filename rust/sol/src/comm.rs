//! Multi-GPU exchange of the path (new: the reference is single-GPU).  One process per GPU, one
//! communicator per Context; NCCL over NVLink inside libsolb.  Bootstrap is the host's business: rank 0 calls
//! `unique_id()` and ships the 128 bytes to the other ranks (MPI, a socket, a file), every rank then calls `Communicator::new`.
use crate::ffi::*;
use crate::{Context, Image2d};
use std::sync::Arc;

pub fn unique_id() -> [u8; SOLB_COMM_ID_BYTES] {
    let mut id = [0u8; SOLB_COMM_ID_BYTES];
    let rc = unsafe { solb_comm_unique_id(id.as_mut_ptr()) };
    if rc != 0 {
        panic!("solb_comm_unique_id: {}", crate::context::last_error(std::ptr::null_mut()));
    }
    id
}

/// Frames of [first, first + n_frames) owned by `rank`: f = first + rank (mod world).
pub fn frames_for_rank(rank: u32, world: u32, n_frames: u32, first: u32) -> Vec<u32> {
    assert!(world >= 1 && rank < world);
    ((first + rank)..(first + n_frames)).step_by(world as usize).collect()
}

/// (tile_row_begin, tile_row_count, tile_row_stride) of `rank` for the tile split with bands of `band_rows` rows.
pub fn tile_rows_for_rank(rank: u32, world: u32, band_rows: u32) -> (u32, u32, u32) {
    (rank * band_rows, band_rows, world * band_rows)
}

pub struct Communicator {
    context: Arc<Context>,
    pub rank: i32,
    pub world: i32,
}

impl Communicator {
    pub fn new(context: Arc<Context>, id: &[u8; SOLB_COMM_ID_BYTES], rank: i32, world: i32) -> Communicator {
        context.check(unsafe { solb_comm_init(context.raw, id.as_ptr(), rank, world) });
        Communicator { context, rank, world }
    }

    /// Frames split: sums the per-rank SOLB_ACCUM_SUM targets onto `root` and resolves there (sum / count, gamma, rgba8).
    pub fn reduce_accum(&self, sum: &Image2d, root: i32, accum_out: Option<&Image2d>, render: Option<&Image2d>) {
        let raw = |t: Option<&Image2d>| t.map(|i| i.raw).unwrap_or(std::ptr::null_mut());
        self.context.check(unsafe { solb_reduce_accum(self.context.raw, sum.raw, root, raw(accum_out), raw(render)) });
    }

    /// Tile split: every rank traced `tile_rows_for_rank(rank, world, band_rows)`; afterwards all hold the whole image.
    pub fn allgather_rows(&self, target: &Image2d, band_rows: u32) {
        self.context.check(unsafe { solb_allgather_rows(self.context.raw, target.raw, band_rows) });
    }

    /// (rank, world, NCCL version)
    pub fn info(&self) -> (i32, i32, i32) {
        let (mut r, mut w, mut v) = (0, 0, 0);
        self.context.check(unsafe { solb_comm_info(self.context.raw, &mut r, &mut w, &mut v) });
        (r, w, v)
    }
}

impl Drop for Communicator {
    fn drop(&mut self) {
        unsafe {
            solb_comm_destroy(self.context.raw);
        }
    }
}
