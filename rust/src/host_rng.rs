//! Seedable construction-time randomness (box heights, sphere centres, Perlin tables).  The reference
//! draws these from the OS-seeded ThreadRng (src/utils.rs:5-15); an explicit xoshiro256++ stream lets
//! the oracle and the GPU see the same scene.  Same algorithm as HostRng in host/rtb/scene.hpp.
use std::cell::RefCell;

pub struct HostRng {
    s: [u64; 4],
}

impl HostRng {
    pub fn new(seed: u64) -> Self {
        let mut z = seed;
        let mut s = [0u64; 4];
        for v in s.iter_mut() {
            z = z.wrapping_add(0x9E3779B97F4A7C15);
            let mut t = z;
            t = (t ^ (t >> 30)).wrapping_mul(0xBF58476D1CE4E5B9);
            t = (t ^ (t >> 27)).wrapping_mul(0x94D049BB133111EB);
            *v = t ^ (t >> 31);
        }
        HostRng { s }
    }
    pub fn next(&mut self) -> u64 {
        let r = (self.s[0].wrapping_add(self.s[3])).rotate_left(23).wrapping_add(self.s[0]);
        let t = self.s[1] << 17;
        self.s[2] ^= self.s[0];
        self.s[3] ^= self.s[1];
        self.s[1] ^= self.s[2];
        self.s[0] ^= self.s[3];
        self.s[2] ^= t;
        self.s[3] = self.s[3].rotate_left(45);
        r
    }
}

thread_local! { static RNG: RefCell<HostRng> = RefCell::new(HostRng::new(20240001)); }

pub fn seed_host_rng(seed: u64) {
    RNG.with(|r| *r.borrow_mut() = HostRng::new(seed));
}
pub fn random_double() -> f64 {
    RNG.with(|r| (r.borrow_mut().next() >> 11) as f64 * (1.0 / 9007199254740992.0))
}
pub fn random_range(min: f64, max: f64) -> f64 {
    min + (max - min) * random_double()
}
pub fn random_int(min: i64, max: i64) -> i64 {
    RNG.with(|r| min + (r.borrow_mut().next() % ((max - min + 1) as u64)) as i64)
}
