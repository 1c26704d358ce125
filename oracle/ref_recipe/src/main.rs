// Prints what rust-bio 1.6.0 itself answers for the cases in tests/golden/ref_cases.tsv (pattern \t text \t k), with
// the ambiguity table hyperex registers (src/utils.rs:251-268) and the call it makes (src/utils.rs:301-309):
//   MyersBuilder::new().ambig(..) x 11 -> build_64(pattern) -> find_all_lazy(text, k).collect()
// One JSON object per line on stdout: {"pattern": "...", "text": "...", "k": 2, "hits": [[end, dist], ...]}
use bio::pattern_matching::myers::MyersBuilder;
use std::io::{self, BufRead};

fn main() {
    let ambigs: [(u8, &[u8]); 11] = [
        (b'M', b"AC"), (b'R', b"AG"), (b'W', b"AT"), (b'S', b"CG"), (b'Y', b"CT"), (b'K', b"GT"),
        (b'V', b"ACGMRS"), (b'H', b"ACTMWY"), (b'D', b"AGTRWK"), (b'B', b"CGTSYK"), (b'N', b"ACGTMRWSYKVHDB"),
    ];
    let mut builder = MyersBuilder::new();
    for (code, equiv) in ambigs.iter() {
        builder.ambig(*code, equiv.iter());
    }
    for line in io::stdin().lock().lines() {
        let line = line.unwrap();
        let f: Vec<&str> = line.split('\t').collect();
        if f.len() != 3 { continue; }
        let (pattern, text, k) = (f[0].as_bytes(), f[1].as_bytes(), f[2].parse::<u8>().unwrap());
        let mut myers = builder.build_64(pattern);
        let hits: Vec<(usize, u8)> = myers.find_all_lazy(text, k).collect();
        let body: Vec<String> = hits.iter().map(|(e, d)| format!("[{},{}]", e, d)).collect();
        println!("{{\"pattern\": \"{}\", \"text\": \"{}\", \"k\": {}, \"hits\": [{}]}}", f[0], f[1], k, body.join(","));
    }
}
